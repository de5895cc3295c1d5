"""CPU tests of the boundary: the C-ABI library builds, loads, exports every symbol the header
declares, and refuses to compute without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "daskr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dkb_[a-z0-9_]+)\s*\(", src)) - {"dkb_res_t", "dkb_jack_t", "dkb_psol_t"})


def test_library_exports_every_header_symbol():
    import daskr_b200
    from daskr_b200.api import EXPORTS

    L = daskr_b200.lib()
    syms = _header_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(EXPORTS) == syms, "api.EXPORTS out of sync with include/daskr_b200.h"
    assert L.dkb_version() == 100


def test_header_cites_reference_interfaces():
    src = open(os.path.join(ROOT, "include", "daskr_b200.h")).read()
    for cite in ("src/daskr.f:1-3", "src/daskr_new.f90:11-21", "src/daskr_rbdpre.f90:102-156",
                 "src/daskr_new.f90:3036-3165", "src/linpack.f90:325-520", "src/daskr_new.f90:828-868"):
        assert cite in src, cite


def test_no_cpu_fallback():
    """Without a CUDA device every entry point that would compute fails loudly."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import daskr_b200 as dk

    with pytest.raises(dk.DaskrError, match="no CPU fallback"):
        dk.init(0)
    with pytest.raises(dk.DaskrError):
        dk.set_option(dk.OPT_FUSED, 1)
    with pytest.raises(dk.DaskrError):
        dk.web_setpar(1, 20, 20)
    info = np.zeros(20, dtype=np.int32)
    info[11] = 1
    with pytest.raises(dk.DaskrError, match="idid=-33"):
        dk.daskr(dk.WEB_RES, 800, 0.0, np.zeros(800), np.zeros(800), 1.0, info, 1e-5, 1e-5, np.zeros(60),
                 np.zeros(840, dtype=np.int32), np.zeros(1), np.zeros(2, dtype=np.int32), dk.WEB_JAC, dk.WEB_PSOL)
    assert info[0] == -1   # DASKR's convention: info(1) = -1 after an illegal-input return


def test_product_does_not_touch_the_oracle():
    """The shipped package must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "daskr_b200")
    for dirpath, _, files in os.walk(pkg):
        if "_build" in dirpath or "_lib" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f)).read()
                if f == "build.py":
                    continue  # build.py compiles the checker for the tests; it never loads it
                for pat in ("liboracle", "oracle_py", "oracle.h", "import oracle", "from oracle", "oracle/_build",
                            "o_daskr", "o_dslvk", "o_dgefa", "o_run_"):
                    assert pat not in txt, (f, pat)
    import subprocess

    out = subprocess.run(["ldd", os.path.join(pkg, "_lib", "libdaskr_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_slab_partition():
    from daskr_b200.slab import partition

    for my in (5, 20, 1024, 4097):
        for n in (1, 2, 3, 4, 8):
            rows = [partition(my, n, r) for r in range(n)]
            assert rows[0][0] == 0
            assert sum(m for _, m in rows) == my
            for (a0, am), (b0, _) in zip(rows, rows[1:]):
                assert a0 + am == b0
            assert max(m for _, m in rows) - min(m for _, m in rows) <= 1
