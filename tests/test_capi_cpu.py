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


def test_block_grouping_tables_and_ownership():
    """Host mirror of dkb_setup_rbgpre (daskr_b200/slab.py) against the oracle's set_grouping restatement, and the
    multi-rank ownership of the group representatives: exactly one owner per group for every slab split."""
    import ctypes as C

    from daskr_b200.slab import partition, rbg_representatives, set_grouping
    from tests import oracle_py

    orc = oracle_py.load()
    for mx, my, nxg, nyg in ((20, 20, 5, 5), (33, 27, 4, 3), (1024, 1024, 5, 5), (17, 90, 1, 7)):
        jigx, jxr = set_grouping(mx, nxg)
        jigy, jyr = set_grouping(my, nyg)
        assert len(jigx) == mx and len(jigy) == my and min(jigx) == 1 and max(jigx) == nxg and max(jigy) == nyg
        assert all(jigx[j - 1] == g + 1 for g, j in enumerate(jxr))      # a representative belongs to its own group
        assert all(jigy[j - 1] == g + 1 for g, j in enumerate(jyr))
        # the oracle (restating src/daskr_rbgpre.f90:160-193) solves every point with the block of group jig:
        # with blocks = group number * identity, psol_rbgpre divides each point by its group number
        ns = 2
        orc.L.o_web_setpar(1, mx, my)
        iw = np.zeros(64, dtype=np.int32)
        orc.L.o_setup_rbgpre.argtypes = [C.c_int] * 7 + [C.c_void_p]
        orc.L.o_setup_rbgpre(mx, my, ns, 1, nxg, nyg, 0, iw.ctypes.data_as(C.c_void_p))
        ngrp = nxg * nyg
        bd = np.concatenate([np.diag([g + 1.0] * ns).ravel() for g in range(ngrp)])
        ip = np.tile(np.arange(1, ns + 1, dtype=np.int32), ngrp)
        b = np.ones(ns * mx * my)
        orc.L.o_psol_rbgpre.argtypes = [C.c_void_p] * 3
        orc.L.o_psol_rbgpre(b.ctypes.data_as(C.c_void_p), bd.ctypes.data_as(C.c_void_p), ip.ctypes.data_as(C.c_void_p))
        grp = np.array([[(jigy[jy] - 1) * nxg + jigx[jx] for jx in range(mx)] for jy in range(my)], dtype=float)
        assert np.allclose(b.reshape(my, mx, ns)[:, :, 0], 1.0 / grp)
        for nranks in (1, 2, 3, 8):
            if my < nranks:
                continue
            owners = np.zeros(ngrp, dtype=int)
            for rank in range(nranks):
                rep = rbg_representatives(mx, my, nxg, nyg, nranks, rank)
                jy0, myloc = partition(my, nranks, rank)
                for g, p in enumerate(rep):
                    if p >= 0:
                        owners[g] += 1
                        jyl, jx = divmod(p, mx)
                        assert jy0 + jyl + 1 == jyr[g // nxg] and jx + 1 == jxr[g % nxg]
            assert (owners == 1).all()


def test_fortran_interface_module_covers_the_header():
    """fortran/daskr_b200.f90 (the ISO_C_BINDING module a Fortran host uses) declares every entry point and callback
    type of include/daskr_b200.h under its C name, and nothing the header does not have."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "daskr_b200.h")).read()
    module = open(os.path.join(root, "fortran", "daskr_b200.f90")).read()
    in_header = set(re.findall(r"\b(dkb_[a-z0-9_]+)\s*\(", header)) | set(re.findall(r"\(\*(dkb_[a-z0-9_]+_t)\)", header))
    bound = set(re.findall(r"name\s*=\s*[\"'](dkb_[a-z0-9_]+)[\"']", module))
    abstract = set(re.findall(r"subroutine\s+(dkb_[a-z0-9_]+_t)\s*\(", module))
    functions = {n for n in in_header if not n.endswith("_t")}
    assert functions - bound == set(), sorted(functions - bound)
    assert bound - functions == set(), sorted(bound - functions)
    callbacks = {n for n in in_header if n.endswith("_t")}
    # the batched rblock type carries a longer name on the Fortran side
    assert callbacks - abstract - {"dkb_rblock_t"} == set(), sorted(callbacks - abstract)
    assert "dkb_rblock_batched_t" in abstract
