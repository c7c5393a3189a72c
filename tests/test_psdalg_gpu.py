"""SURVEY 8f row 2 (first part): vecsym, sqrtinv and qrK plugins against the reference's own MEX files.
Gates: vecsym exact, sqrtinv 1e-14, qrK 1e-10 (q in product form and triu(r); the reference documents tril(r,-1)
as undefined, qrK.c:80-82)."""
import numpy as np
import pytest

from helpers import gpu, ref, relerr
from sedumi_b200.host import cones

pytestmark = pytest.mark.gpu


def _K(l, q, sreal, sherm=()):
    K = cones.finish_K({"l": float(l), "q": np.array(q, dtype=float), "s": np.array(list(sreal) + list(sherm), dtype=float),
                        "rsdpN": len(sreal)})
    return K, cones.K_for_mex(K)


@pytest.mark.parametrize("sreal,sherm", [((5, 3), ()), ((40,), ()), ((4,), (3, 6)), ((), (5,))])
def test_vecsym(sreal, sherm):
    K, Km = _K(3, (4,), sreal, sherm)
    rng = np.random.default_rng(1)
    n = 3 + 4 + sum(k * k for k in sreal) + 2 * sum(k * k for k in sherm)
    x = rng.standard_normal(n)
    assert np.array_equal(gpu.vecsym(x, Km), ref.vecsym(x, Km))


@pytest.mark.parametrize("sreal,sherm", [((5, 3), ()), ((70, 35), ()), ((4,), (3, 6))])
def test_sqrtinv(sreal, sherm):
    K, Km = _K(2, (3,), sreal, sherm)
    rng = np.random.default_rng(2)
    lenud = sum(k * k for k in sreal) + 2 * sum(k * k for k in sherm)
    q = rng.standard_normal(lenud)
    v = np.exp(rng.standard_normal(2 + 2 * 1 + sum(sreal) + sum(sherm)))
    assert relerr(gpu.sqrtinv(q, v, Km), ref.sqrtinv(q, v, Km)) <= 1e-14


@pytest.mark.parametrize("s", [(1,), (2,), (7, 4), (33, 64, 5), (200,)])
def test_qrK(s):
    K, Km = _K(1, (), s)
    rng = np.random.default_rng(sum(s))
    x = rng.standard_normal(sum(k * k for k in s))
    if len(s) > 1:
        x[:s[0] * s[0]].reshape(s[0], s[0], order="F")[:, 1] = 0.0         # an all-zero column: beta = 1 (qrK.c:103-104)
    qg, rg = gpu.qrK(x, Km, nlhs=2)
    qr_, rr = ref.qrK(x, Km, nlhs=2)
    assert relerr(qg, qr_) <= 1e-10
    off = 0
    for n in s:
        Rg = np.triu(rg.ravel()[off:off + n * n].reshape(n, n, order="F"))
        Rr = np.triu(rr.ravel()[off:off + n * n].reshape(n, n, order="F"))
        assert relerr(Rg, Rr) <= 1e-10
        # and it is a QR factorisation: |R'R - X'X| small
        X = x[off:off + n * n].reshape(n, n, order="F")
        assert relerr(Rg.T @ Rg, X.T @ X) <= 1e-10
        off += n * n


def test_qrK_frames_feed_psdframeit():
    """qrK's q is the product form psdframeit consumes (wregion.m / frameit.m): Q diag(lab) Q' from our qrK through our
    psdframeit equals the same through the reference's."""
    s = (12, 9)
    K, Km = _K(1, (), s)
    rng = np.random.default_rng(5)
    x = rng.standard_normal(sum(k * k for k in s))
    lab = np.exp(rng.standard_normal(sum(s)))
    qg, _ = gpu.qrK(x, Km, nlhs=2)
    qr_, _ = ref.qrK(x, Km, nlhs=2)
    assert relerr(gpu.psdframeit(lab, qg, Km), ref.psdframeit(lab, qr_, Km)) <= 1e-10


# ---- the M-only pieces (psdjmul.m, triumtriu.m, psdfactor.m, psdinvscale.m): numpy restatements as the oracle
# ("parity unpinned" beyond that: the reference cannot run M code here)
import os
import sys

from helpers import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import restate  # noqa: E402


@pytest.mark.parametrize("s", [(3,), (40, 7), (130,), (200, 65)])
def test_psdjmul_triumtriu(s):
    K, Km = _K(2, (), s)
    rng = np.random.default_rng(sum(s))
    lenud = sum(k * k for k in s)
    x, y = rng.standard_normal(2 + lenud), rng.standard_normal(2 + lenud)        # PSD part = tail
    assert relerr(np.ravel(gpu.psdjmul(x, y, Km)), restate.psdjmul(x, y, K)) <= 1e-12
    assert relerr(np.ravel(gpu.triumtriu(x, y, Km)), restate.triumtriu(x, y, K)) <= 1e-12


@pytest.mark.parametrize("s", [(4,), (33, 9), (161,)])
def test_psdfactor_and_psdinvscale(s):
    K, Km = _K(1, (), s)
    rng = np.random.default_rng(7 + sum(s))
    blocks = []
    for n in s:
        G = rng.standard_normal((n, n + 3))
        blocks.append(G @ G.T / n + 0.1 * np.eye(n))
    x = np.concatenate([B.ravel(order="F") for B in blocks])
    ug, pos = gpu.psdfactor(x, Km, nlhs=2)
    ur, posr = restate.psdfactor(x, K)
    assert float(np.asarray(pos).ravel()[0]) == 1.0 and posr
    assert relerr(np.ravel(ug), ur) <= 1e-10
    # the factor is the `ud` psdinvscale takes (upper triangle = L'): Y = T \ (X / T') with X symmetric
    X = np.concatenate([(lambda M: (M + M.T).ravel(order="F"))(rng.standard_normal((n, n))) for n in s])
    assert relerr(np.ravel(gpu.psdinvscale(ug.ravel(), X, Km)), restate.psdinvscale(ur, X, K)) <= 1e-9
    # an indefinite block: flag, and nothing from that block on
    bad = x.copy()
    if len(s) > 1:
        o = s[0] * s[0]
        bad[o] = -1.0                                       # (1,1) entry of the second block
        ub, posb = gpu.psdfactor(bad, Km, nlhs=2)
        urb, posrb = restate.psdfactor(bad, K)
        assert float(np.asarray(posb).ravel()[0]) == 0.0 and not posrb
        assert relerr(ub.ravel()[:o], urb[:o]) <= 1e-10 and not np.any(ub.ravel()[o:])


@pytest.mark.parametrize("s", [(1,), (2, 3), (40, 7), (129,), (200, 65)])
def test_psdeig_minpsdeig(s):
    """psdeig.m / minpsdeig.m (M-only; the reference calls the host's eig): eigenvalues against LAPACK, eigenvectors through
    what makes them unique -- orthonormality and Q diag(lab) Q' = (X + X')/2."""
    K, Km = _K(1, (), s)
    rng = np.random.default_rng(31 + sum(s))
    lenud = sum(k * k for k in s)
    x = rng.standard_normal(1 + lenud)
    lab_r = restate.psdeig(x, K)
    lab_g, q_g = gpu.psdeig(x, Km, nlhs=2)
    lab_g, q_g = np.ravel(lab_g), np.ravel(q_g)
    scale = max(1.0, np.abs(lab_r).max())
    assert np.abs(lab_g - lab_r).max() <= 1e-10 * scale
    assert np.abs(np.ravel(gpu.psdeig(x, Km)) - lab_r).max() <= 1e-10 * scale          # values only
    o = lo = 0
    for n, X in zip(s, restate._blocks(x, K)):
        Q = q_g[o:o + n * n].reshape(n, n, order="F")
        lam = lab_g[lo:lo + n]
        assert np.all(np.diff(lam) >= 0)
        assert np.abs(Q.T @ Q - np.eye(n)).max() <= 1e-12 * n
        assert np.abs(Q @ np.diag(lam) @ Q.T - 0.5 * (X + X.T)).max() <= 1e-10 * scale
        o += n * n
        lo += n
    assert abs(float(np.ravel(gpu.minpsdeig(x, Km))[0]) - restate.minpsdeig(x, K)) <= 1e-10 * scale


def test_psdeig_repeated_and_zero_eigenvalues():
    s = (6, 5)
    K, Km = _K(0, (), s)
    A = np.ones((6, 6))                                    # eigenvalues 0 (x5), 6
    B = np.diag([1.0, 1.0, 2.0, 2.0, 3.0])
    x = np.concatenate([A.ravel(order="F"), B.ravel(order="F")])
    lab, q = gpu.psdeig(x, Km, nlhs=2)
    lab, q = np.ravel(lab), np.ravel(q)
    assert np.abs(lab - np.r_[0, 0, 0, 0, 0, 6, 1, 1, 2, 2, 3]).max() <= 1e-12
    Q = q[:36].reshape(6, 6, order="F")
    assert np.abs(Q @ np.diag(lab[:6]) @ Q.T - A).max() <= 1e-12

