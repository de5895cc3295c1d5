"""examples/example_web.cpp: the reference's demonstration program (example/example_web.f90:721-990) as a compiled
host program on the C ABI.  CPU: it compiles, links against the library and refuses to run without a GPU.
GPU: it reproduces the table of the Python host (same DASKR calls) for METH = 1 (rbdpre) and METH = 2 (rbgpre)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "example_web")
    lib = os.path.join(ROOT, "daskr_b200", "_lib")
    cmd = ["g++", "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "example_web.cpp"),
           "-L" + lib, "-ldaskr_b200", "-Wl,-rpath," + lib, "-o", exe]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert p.returncode == 0, p.stdout
    return exe


def test_example_compiles_and_needs_a_gpu(tmp_path):
    import torch

    exe = _build(tmp_path)
    if not torch.cuda.is_available():
        p = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
        assert p.returncode != 0 and "no CPU fallback" in p.stdout


@pytest.mark.gpu
def test_example_matches_the_python_host(dk, tmp_path):
    exe = _build(tmp_path)
    dk.finalize()                                   # the program brings up its own context
    p = subprocess.run([exe, "20", "20", "1"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    dk.init(0)
    assert p.returncode == 0, p.stdout[-2000:]
    runs = p.stdout.split("Linear solver method flag")[1:]
    assert len(runs) == 2
    touts = [1e-8, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0, 8.0, 9.0, 10.0]
    for text, ng in zip(runs, (0, 5)):
        assert text.count("Root found") == 1
        rows = [ln.split() for ln in text.splitlines() if re.match(r"^\s*\d\.\d{5}e[-+]\d\d\s", ln)]
        table = np.array([[float(x) for x in r] for r in rows])     # t <c1> NSTEP NRE NNI NLI NPE NQ H AVLIN
        g = dk.solve_web(1, 20, 20, touts, jpre=3, roots=True, nxg=ng, nyg=ng)
        # rows: IC call, then one per output time plus one for the root
        out_rows = table[[i for i in range(len(table)) if any(abs(table[i, 0] - t) <= 1e-12 * t for t in touts)][-len(touts):]]
        assert np.array_equal(out_rows[:, 2].astype(int), g["counters"][:, 10])       # NSTEP
        assert np.array_equal(out_rows[:, 5].astype(int), g["counters"][:, 19])       # NLI
        c1 = g["y"].reshape(len(touts), 20, 20, 2)[:, :, :, 0].mean(axis=(1, 2))
        assert np.allclose(out_rows[:, 1], c1, rtol=0, atol=6e-6)                      # printed with 5 decimals
