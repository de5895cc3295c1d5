"""Multi-GPU parity inside the driver-run suite: tests/mgpu_check.py (y-slabs, NCCL halo rows + scalar allreduce, the
overlapped-slab Gauss-Seidel factor, block grouping, roots, state export/import; every case against the oracle's
full-mesh run) launched under torchrun at every world size the box offers.  Skips itself below two devices."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_mgpu_check(world):
    n = _device_count()
    if n < world:
        pytest.skip("needs %d GPUs, this box has %d" % (world, n))
    port = 29500 + world * 7 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")]
    env = dict(os.environ)
    env.pop("NCCL_DEBUG", None)
    p = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    lines = [ln for ln in p.stdout.splitlines() if "->" in ln or "world=" in ln]
    print("\n".join(lines))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "mgpu_world%d.log" % world), "w") as f:
            f.write(p.stdout)
    assert p.returncode == 0, p.stdout[-4000:]
    assert lines and not any("FAIL" in ln for ln in lines)
