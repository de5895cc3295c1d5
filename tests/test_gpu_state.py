"""Checkpoint-style continuation: export the device-side state + the caller's rwork / iwork,
tear the library down, bring it up again, import, continue -- the continued run must take exactly the steps of
the uninterrupted one (bitwise equal outputs and counters)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NP, MX, MY = 1, 24, 20
NS = 2 * NP
TOUTS = [1e-6, 1e-4, 1e-3, 1e-2, 1e-1]


def _setup(dk):
    probe = np.zeros(64, dtype=np.int32)
    dk.setup_rbdpre(MX, MY, NS, NP, 0, probe)
    neq = NS * MX * MY
    iwork = np.zeros(40 + neq, dtype=np.int32)
    dk.setup_rbdpre(MX, MY, NS, NP, 40, iwork)
    dk.web_setpar(NP, MX, MY)
    return neq, iwork


def _run(dk, jpre, break_after=None):
    neq, iwork = _setup(dk)
    info = np.zeros(20, dtype=np.int32)
    info[[10, 11, 13, 14, 15]] = 1
    rwork, rpar, ipar = np.zeros(60), np.zeros(1), np.array([jpre, 0], dtype=np.int32)
    c, cdot = dk.web_cinit(neq)
    args = lambda: (rwork, iwork, rpar, ipar, dk.WEB_JAC, dk.WEB_PSOL)  # noqa: E731
    t, idid = dk.daskr(dk.WEB_RES, neq, 0.0, c, cdot, TOUTS[0], info, 1e-5, 1e-5, *args())
    assert idid == 4
    info[10] = 0
    ys, cnts = [], []
    for k, tout in enumerate(TOUTS):
        if break_after is not None and k == break_after:
            # ---- checkpoint: device state + the caller's arrays ----
            saved = dict(state=dk.state_export(), rwork=rwork.copy(), iwork=iwork.copy(), info=info.copy(), t=t,
                         c=c.copy(), cdot=cdot.copy())
            dk.finalize()
            dk.init(0)
            neq, _ = _setup(dk)                       # a fresh process would do exactly this
            dk.state_import(saved["state"])
            rwork, iwork, info = saved["rwork"], saved["iwork"], saved["info"]
            c, cdot, t = saved["c"], saved["cdot"], saved["t"]
        t, idid = dk.daskr(dk.WEB_RES, neq, t, c, cdot, tout, info, 1e-5, 1e-5, *args())
        assert idid == 3
        ys.append(c.copy())
        cnts.append(iwork[:40].copy())
    return np.array(ys), np.array(cnts)


@pytest.mark.parametrize("jpre", [1, 3])
def test_export_import_continues_bitwise(dk, jpre):
    y0, c0 = _run(dk, jpre)
    y1, c1 = _run(dk, jpre, break_after=2)
    assert np.array_equal(c0, c1)
    assert np.array_equal(y0, y1)


def test_import_rejects_garbage(dk):
    _setup(dk)
    with pytest.raises(dk.DaskrError):
        dk.state_import(np.zeros(4096, dtype=np.uint8))
