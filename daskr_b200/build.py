"""Build helpers: compile the CUDA library (sm_100a) and the CPU oracle.

Both builds are plain `make` invocations of committed Makefiles; nvcc
cross-compiles for sm_100a without a GPU.  The oracle is TEST INFRASTRUCTURE
(see oracle/oracle.h) -- building it here does not make the product use it.
"""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "daskr_b200", "csrc")
LIB = os.path.join(ROOT, "daskr_b200", "_lib", "libdaskr_b200.so")
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "liboracle.so")


def _run(cmd, cwd):
    p = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), p.stdout[-4000:]))
    return p.stdout


def build_cuda(jobs=8):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... -> daskr_b200/_lib/libdaskr_b200.so"""
    _run(["make", "-j%d" % jobs], CSRC)
    if not os.path.exists(LIB):
        raise RuntimeError("make succeeded but %s is missing" % LIB)
    return LIB


def build_oracle():
    """g++ -O2 -ffp-contract=off -> oracle/_build/liboracle.so (the checker, never the product)."""
    _run(["make"], ORACLE_DIR)
    if not os.path.exists(ORACLE_LIB):
        raise RuntimeError("make succeeded but %s is missing" % ORACLE_LIB)
    return ORACLE_LIB


if __name__ == "__main__":
    print(build_cuda())
    print(build_oracle())
