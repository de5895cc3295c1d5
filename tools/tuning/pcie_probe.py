"""Host<->device copy rates on this box (pinned torch tensors, CUDA events): context for the e2e leg of bench.py."""
import time, torch
torch.cuda.init()
dev = torch.device("cuda", 0)
s2 = torch.cuda.Stream()
for mb in (64, 335, 1024, 2684):
    n = mb * (1 << 20) // 8
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    h2 = torch.empty(n, dtype=torch.float64).pin_memory()
    d = torch.zeros(n, dtype=torch.float64, device=dev)
    d2 = torch.zeros(n, dtype=torch.float64, device=dev)
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True)),
                     ("d2h x2 back to back", lambda: (h.copy_(d, non_blocking=True), h2.copy_(d2, non_blocking=True)))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        nbytes = n * 8 * (2 if "x2" in name else 1)
        print("%5d MB %-20s %8.2f ms  %6.1f GB/s" % (mb, name, ms, nbytes / ms / 1e6), flush=True)
    # d2h on a side stream while a bandwidth-bound kernel runs on the main stream
    big = torch.zeros(1 << 28, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s2):
        e0.record(s2)
        h.copy_(d, non_blocking=True)
        e1.record(s2)
    for _ in range(20): big.add_(1.0)
    torch.cuda.synchronize()
    print("%5d MB d2h under a streaming kernel %8.2f ms  %6.1f GB/s" % (mb, e0.elapsed_time(e1), n * 8 / e0.elapsed_time(e1) / 1e6), flush=True)
    del h, h2, d, d2, big
