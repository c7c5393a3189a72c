"""Parity of adendotd / adenscale (dense Lorentz blocks, getDAtm.m:45 / deninfac.m:62) on synthetic
inputs with the reference's layout: dense.cols = [LP dense cols, dense-block trace rows, dense
norm-bound columns]."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import gpu, ref, relerr

pytestmark = pytest.mark.gpu


def _inputs(seed, m=40):
    rng = np.random.default_rng(seed)
    Kl, qs = 3, [4, 6, 3, 5, 7]
    lorN = len(qs)
    first = Kl + 1 + lorN                       # 1-based index of the first norm-bound row (x0 + LP + traces)
    blkstart = np.cumsum(np.r_[first, np.array(qs) - 1]).astype(float)
    dq = np.array([2, 4])                       # dense Lorentz blocks (1-based)
    nl = 1
    # dense norm-bound columns: a few rows inside the dense blocks, ascending
    dencols = []
    for k in dq:
        lo, hi = int(blkstart[k - 1]), int(blkstart[k])
        dencols += sorted(rng.choice(np.arange(lo, hi), size=min(2, hi - lo), replace=False).tolist())
    cols = np.r_[2.0, Kl + 1 + dq, dencols]     # LP col, trace rows of the dense blocks, norm-bound cols
    ncol = nl + len(dq) + len(dencols)
    A = sp.random(m, ncol, density=0.4, random_state=np.random.RandomState(seed), format="csc")
    dense = {"l": float(nl), "q": dq.astype(float).reshape(-1, 1), "cols": cols.reshape(-1, 1), "A": A}
    d = {"q1": rng.standard_normal(lorN), "q2": rng.standard_normal(int(sum(qs)) - lorN),
         "det": np.exp(rng.standard_normal(lorN))}
    adotd = sp.random(m, len(dq), density=0.3, random_state=np.random.RandomState(seed + 1), format="csc")
    pat = sp.csc_matrix((adotd != 0).astype(float))
    j = 0
    for kk, k in enumerate(dq):
        col = (A[:, nl + kk] != 0).astype(float)
        while j < len(dencols) and dencols[j] < blkstart[k]:
            col = col + (A[:, nl + len(dq) + j] != 0).astype(float)
            j += 1
        pat[:, kk] = pat[:, kk] + col
    Ablk = sp.csc_matrix(pat != 0, dtype=float)
    return dense, d, adotd, Ablk, blkstart.reshape(1, -1)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_adendotd_adenscale(seed):
    dense, d, adotd, Ablk, bs = _inputs(seed)
    r = ref.adendotd(dense, d, adotd, Ablk, bs)
    g = gpu.adendotd(dense, d, adotd, Ablk, bs)
    assert np.array_equal(r.indices, g.indices) and relerr(g.data, r.data) <= 1e-12
    assert np.array_equal(gpu.adenscale(dense, d, bs), ref.adenscale(dense, d, bs))


def test_no_dense_blocks():
    m = 10
    dense = {"l": 0.0, "q": np.zeros((0, 1)), "cols": np.zeros((0, 1)), "A": sp.csc_matrix((m, 0))}
    d = {"q1": np.ones(2), "q2": np.ones(4), "det": np.ones(2)}
    bs = np.array([[4.0, 6.0, 8.0]])
    out = gpu.adendotd(dense, d, sp.csc_matrix((m, 0)), sp.csc_matrix((m, 0)), bs)
    assert out.shape == (m, 0)
    assert gpu.adenscale(dense, d, bs).size == 0
