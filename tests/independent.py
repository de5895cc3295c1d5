"""An integrator for the food-web DAE M c' = f(c) (M = 1 on the prey, 0 on the predators) that shares nothing with
DASKR but the right-hand side: implicit Euler with a dense finite-difference Jacobian and numpy's LU on geometric time
grids, Richardson-extrapolated twice.  Test infrastructure (tests/test_oracle_cpu.py pins the oracle with it,
tests/test_gpu_parity.py the CUDA path)."""
import numpy as np


def extrapolated_implicit_euler(f, M, c0, t0, touts, nper=(100, 200, 400)):
    """Solution at touts (all > t0) from three grids with nper[i] steps per output interval (each twice the one before);
    O(h) and O(h^2) error terms removed."""
    neq = c0.size

    def jac(c):
        fc, J = f(c), np.empty((neq, neq))
        for j in range(neq):
            dc = 1e-7 * max(abs(c[j]), 1e-3)
            cp = c.copy()
            cp[j] += dc
            J[:, j] = (f(cp) - fc) / dc
        return J

    def step(c, h):
        cn, J = c.copy(), None
        for it in range(20):
            if J is None or it == 3:
                J = jac(cn)
            d = np.linalg.solve(np.diag(M / h) - J, -(M * (cn - c) / h - f(cn)))
            cn += d
            if np.max(np.abs(d) / (np.abs(cn) + 1e-300)) < 1e-13:
                return cn
        raise AssertionError("implicit Euler: Newton did not converge")

    def integrate(n):
        out, c, t = [], c0.copy(), t0
        for t1 in touts:
            grid = np.exp(np.linspace(np.log(t), np.log(t1), n + 1))
            for a, b in zip(grid[:-1], grid[1:]):
                c = step(c, b - a)
            t = t1
            out.append(c.copy())
        return np.array(out)

    A, B, Cc = (integrate(n) for n in nper)
    return (4 * (2 * Cc - B) - (2 * B - A)) / 3
