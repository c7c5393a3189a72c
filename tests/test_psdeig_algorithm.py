"""CPU statement of the algorithm inside `alg_jacobi_kernel` (sedumi_b200/csrc/psdalg.cu): cyclic two-sided Jacobi in the
round-robin ordering, all 2 x 2 sub-blocks of a step rotated from the step's starting matrix.  The GPU tests compare the
kernel with LAPACK; this one pins the pieces that are easy to get wrong -- the tournament schedule (every index pair exactly
once per sweep, the dummy index of odd orders), the rotation formula and the stopping rule -- without a GPU."""
import numpy as np
import pytest


def schedule(m, st):
    """Index pairs of step `st` for an even number of players m (player m-1 stays, the others rotate)."""
    out = [(m - 1, st % (m - 1))]
    for k in range(1, m // 2):
        out.append(((st + k) % (m - 1), (st - k + (m - 1)) % (m - 1)))
    return [(min(a, b), max(a, b)) for a, b in out]


def jacobi_eig(X, max_sweeps=30):
    n = X.shape[0]
    A = 0.5 * (X + X.T)
    V = np.eye(n)
    if n == 1:
        return np.diag(A).copy(), V, 0
    m = n + (n & 1)
    fro2 = float((A * A).sum())
    sweeps = 0
    for sweeps in range(max_sweeps + 1):
        off2 = float(((A - np.diag(np.diag(A))) ** 2).sum())          # summed directly, never as fro2 - diag2
        if not off2 > 1e-29 * fro2 or sweeps == max_sweeps:
            break
        for st in range(m - 1):
            J = np.eye(n)
            for p, q in schedule(m, st):
                if q >= n or A[p, q] == 0.0:
                    continue
                tau = (A[q, q] - A[p, p]) / (2.0 * A[p, q])
                t = 1.0 if tau == 0.0 else np.copysign(1.0, tau) / (abs(tau) + np.sqrt(1.0 + tau * tau))
                c = 1.0 / np.sqrt(1.0 + t * t)
                J[p, p] = c; J[q, q] = c; J[p, q] = t * c; J[q, p] = -t * c
            A = J.T @ A @ J                                           # = every 2 x 2 sub-block J_K' A_KL J_L
            V = V @ J
    lam = np.diag(A).copy()
    order = np.argsort(lam, kind="stable")
    return lam[order], V[:, order], sweeps


@pytest.mark.parametrize("m", [2, 4, 8, 10, 66])
def test_schedule_meets_every_pair_once_per_sweep(m):
    seen = set()
    for st in range(m - 1):
        ps = schedule(m, st)
        idx = sorted(i for pq in ps for i in pq)
        assert idx == list(range(m))                                  # a step touches every index exactly once
        seen.update(ps)
    assert len(seen) == m * (m - 1) // 2


@pytest.mark.parametrize("n", [1, 2, 3, 7, 40, 65])
def test_converges_to_lapack(n):
    rng = np.random.default_rng(100 + n)
    X = rng.standard_normal((n, n))
    lam, Q, sweeps = jacobi_eig(X)
    S = 0.5 * (X + X.T)
    w = np.linalg.eigvalsh(S)
    assert sweeps <= 10
    assert np.abs(lam - w).max() <= 1e-12 * max(1.0, np.abs(w).max())
    assert np.abs(Q.T @ Q - np.eye(n)).max() <= 1e-13 * max(n, 1)
    assert np.abs(Q @ np.diag(lam) @ Q.T - S).max() <= 1e-12 * max(1.0, np.abs(w).max())


def test_repeated_zero_and_diagonal_inputs():
    lam, Q, sweeps = jacobi_eig(np.ones((6, 6)))
    assert np.abs(lam - np.r_[0, 0, 0, 0, 0, 6.0]).max() <= 1e-14 and sweeps <= 2
    lam, _, sweeps = jacobi_eig(np.diag([3.0, 1.0, 2.0, 1.0]))
    assert np.array_equal(lam, [1.0, 1.0, 2.0, 3.0]) and sweeps == 0      # nothing to rotate
    lam, _, sweeps = jacobi_eig(np.zeros((5, 5)))
    assert not lam.any() and sweeps == 0
