"""The reference's own test problems (test/test_krdem1.f90, test/test_krdem2.f90) through the GPU library's
dkb_daskr with user callbacks (device arrays), Krylov option: same acceptance values as the reference's tests,
and the same step / residual / root-function counts as the oracle."""
import ctypes as C

import numpy as np
import pytest

from tests import krdem
from tests.test_oracle_krdem import _oracle_call

pytestmark = pytest.mark.gpu


def _gpu_call(dk):
    rpar, ipar = np.zeros(4), np.zeros(4, dtype=np.int32)

    def call(res, neq, t, y, yd, tout, info, rtol, atol, rwork, iwork, jac, psol, rt, nrt, jroot):
        return dk.daskr(res, neq, t, y, yd, tout, info, rtol, atol, rwork, iwork, rpar, ipar, jac, psol, rt, nrt, jroot)

    return call


def test_krdem1_gpu(dk, oracle):
    dk.finalize()
    dk.init(0)                          # no mesh / preconditioner set-up: a plain user problem
    L = dk.lib()
    L.dkb_memcpy_d2h.restype = C.c_int
    ero, roots, iwork = krdem.run_krdem1(_gpu_call(dk), krdem.DeviceArrays(L))
    krdem.check_krdem1(ero, roots)
    ero_o, roots_o, iwork_o = krdem.run_krdem1(_oracle_call(oracle.L), krdem.HostArrays())
    assert [(round(t, 9), j.tolist()) for t, j in roots] == [(round(t, 9), j.tolist()) for t, j in roots_o]
    assert np.array_equal(iwork[[10, 11, 13, 14, 18, 19, 35]], iwork_o[[10, 11, 13, 14, 18, 19, 35]])


def test_krdem2_gpu(dk, oracle):
    dk.finalize()
    dk.init(0)
    roots, iwork, y = krdem.run_krdem2(_gpu_call(dk), krdem.DeviceArrays(dk.lib()))
    krdem.check_krdem2(roots)
    roots_o, iwork_o, y_o = krdem.run_krdem2(_oracle_call(oracle.L), krdem.HostArrays())
    assert len(roots) == len(roots_o)
    for (t, j), (to, jo) in zip(roots, roots_o):
        assert j == jo and abs(t - to) < 1e-3 * to
    assert abs(int(iwork[10]) - int(iwork_o[10])) <= 0.05 * iwork_o[10]
