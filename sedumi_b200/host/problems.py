"""Workloads: the reference's example problems (as converted fixtures) and the synthetic
cone problems named in BASELINE.json `configs`, plus NT-scaling generators.

Generators follow SURVEY.md section 8(d) ("Concrete inputs"); the random-problem recipe
mirrors the spirit of conversion/feasreal.m:43-93 (b = A*vec(I), c = vec(I) + A'*y0).
All are deterministic in their seed.
"""
from __future__ import annotations

import os

import numpy as np
import scipy.sparse as sp

from .cones import pretransfo, psd_dims

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))),
                       "tests", "golden", "problems")
SEED0 = 20260924


def load_fixture(name: str):
    """Raw (At, b, c, K) of a reference example, from tests/golden/problems/<name>.npz
    (converted from examples/<name>.mat by tests/golden/make_problem_fixtures.py)."""
    z = np.load(os.path.join(_GOLDEN, name + ".npz"))
    At = sp.csc_matrix((z["At_data"], z["At_indices"], z["At_indptr"]), shape=tuple(z["At_shape"]))
    K = {k[2:]: z[k] for k in z.files if k.startswith("K_")}
    return At, z["b"], z["c"], K


def internal_problem(name: str):
    At, b, c, K = load_fixture(name)
    return pretransfo(At, b, c, K)[:4]


# --------------------------------------------------------------------------- synthetic
def _sprandsym_lower(n, nnz_target, rng):
    """Random symmetric sparse n x n (values N(0,1)); returns (rows, cols, vals) of the FULL matrix."""
    k = max(nnz_target // 2, 1)
    r = rng.integers(0, n, size=k)
    c = rng.integers(0, n, size=k)
    v = rng.standard_normal(k)
    lo = np.maximum(r, c)
    hi = np.minimum(r, c)
    key = lo * n + hi
    _, first = np.unique(key, return_index=True)
    lo, hi, v = lo[first], hi[first], v[first]
    off = lo != hi
    rows = np.r_[lo, hi[off]]
    cols = np.r_[hi, lo[off]]
    vals = np.r_[v, v[off]]
    return rows, cols, vals


def synth_blockdiag_sdp(nblk=64, n=200, m=5000, nlink=72, density=0.02, dense_lp=0, seed=SEED0 + 4):
    """Config 4: block-diagonal SDP, `nblk` PSD blocks of order `n`, `m` constraints.
    Constraint j < m-nlink is local to block (j mod nblk) with a sprandsym(n, density)
    coefficient; the last `nlink` constraints are random diagonals on every block
    (arrow-shaped ADA: nblk independent etree subtrees + a border).  `dense_lp` > 0 adds
    that many LP variables appearing in every constraint (config 4'': exercises the
    dense-column path dpr1fact/fwdpr1/bwdpr1)."""
    rng = np.random.default_rng(seed)
    nn = n * n
    N = dense_lp + nblk * nn
    rows, cols, vals = [], [], []
    nloc = m - nlink
    for j in range(nloc):
        k = j % nblk
        r, c, v = _sprandsym_lower(n, int(density * nn), rng)
        rows.append(dense_lp + k * nn + c * n + r)
        cols.append(np.full(r.size, j))
        vals.append(v)
    for j in range(nloc, m):
        for k in range(nblk):
            dvals = rng.standard_normal(n)
            idx = np.arange(n)
            rows.append(dense_lp + k * nn + idx * n + idx)
            cols.append(np.full(n, j))
            vals.append(dvals)
    if dense_lp:
        for i in range(dense_lp):
            rows.append(np.full(m, i))
            cols.append(np.arange(m))
            vals.append(rng.standard_normal(m))
    At = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, m))
    K = {"l": dense_lp, "s": np.full(nblk, n)}
    x0 = np.r_[np.ones(dense_lp), np.tile(np.eye(n).ravel(), nblk)]
    b = np.asarray(At.T @ x0).ravel()
    c = x0 + np.asarray(At @ rng.standard_normal(m)).ravel()
    return At, b, c, K


def synth_maxcut(n=4000, p=0.005, seed=SEED0 + 5):
    """Config 5: MaxCut relaxation, one PSD block of order n, A_j = e_j e_j', b = 1, c = -Lap/4."""
    rng = np.random.default_rng(seed)
    nedge = int(p * n * (n - 1) / 2)
    i = rng.integers(0, n, nedge)
    j = rng.integers(0, n, nedge)
    keep = i != j
    i, j = i[keep], j[keep]
    W = sp.coo_matrix((np.ones(i.size), (i, j)), shape=(n, n))
    W = sp.csr_matrix(((W + W.T) > 0).astype(np.float64))
    Lap = sp.diags(np.asarray(W.sum(axis=1)).ravel()) - W
    idx = np.arange(n)
    At = sp.csc_matrix((np.ones(n), (idx * n + idx, idx)), shape=(n * n, n))
    b = np.ones(n)
    c = -np.asarray(Lap.todense()).ravel(order="F") / 4.0
    return At, b, c, {"l": 0, "s": np.array([n])}


def synth_dense_sdp(n=1000, m=4000, rank=8, seed=SEED0 + 6):
    """Config 5': one PSD block of order n, m constraints with DENSE symmetric
    coefficients A_j = sum_{r<rank} v v' (a real contraction for getada3)."""
    rng = np.random.default_rng(seed)
    cols = []
    for j in range(m):
        V = rng.standard_normal((n, rank)) / np.sqrt(n)
        cols.append((V @ V.T).ravel(order="F"))
    At = sp.csc_matrix(np.column_stack(cols))
    x0 = np.eye(n).ravel()
    b = np.asarray(At.T @ x0).ravel()
    c = x0 + np.asarray(At @ rng.standard_normal(m)).ravel()
    return At, b, c, {"l": 0, "s": np.array([n])}


def synth_small_mixed(seed=SEED0, m=40, l=6, q=(4, 3, 5), s=(7, 5, 4), density=0.3, f=0, r=()):
    """Small LP + Lorentz + PSD problem for fast parity tests."""
    rng = np.random.default_rng(seed)
    N = f + l + sum(q) + sum(r) + sum(k * k for k in s)
    cols = []
    for j in range(m):
        parts = [rng.standard_normal(f + l + sum(q) + sum(r)) * (rng.random(f + l + sum(q) + sum(r)) < density)]
        for k in s:
            M = rng.standard_normal((k, k)) * (rng.random((k, k)) < density)
            M = M + M.T
            parts.append(M.ravel(order="F"))
        cols.append(np.concatenate(parts))
    At = sp.csc_matrix(np.column_stack(cols))
    x0 = np.concatenate([np.zeros(f), np.ones(l)] + [np.r_[2.0, np.zeros(k - 1)] for k in q] +
                        [np.r_[1.0, 1.0, np.zeros(k - 2)] for k in r] +
                        [np.eye(k).ravel() for k in s])
    b = np.asarray(At.T @ x0).ravel()
    c = x0 + np.asarray(At @ rng.standard_normal(m)).ravel()
    K = {"f": f, "l": l, "q": np.array(q), "r": np.array(r), "s": np.array(s)}
    return At, b, c, K


# --------------------------------------------------------------------------- scalings
def scaling(K: dict, kind: str = "S1", seed: int = SEED0, base: dict | None = None) -> dict:
    """NT scaling ``d`` on an internal-form cone K:
      S0  iteration-0 scaling is produced by setup.sdinit_scaling (not here)
      S1  "mid-run": U_k = triu(qr(I + 0.3 G).R) with positive diagonal, d.l ~ exp(N(0,1)),
          Lorentz d.det ~ exp(N(0,1)), d.q2 ~ N(0,1), d.q1 = sqrt(det + |q2|^2)
      S2  "late": eigenvalues of U_k'U_k log-uniform in [1e-6, 1e6]
    """
    rng = np.random.default_rng(seed)
    nl = int(K["l"])
    q = np.asarray(K["q"], dtype=np.int64)
    s = np.asarray(K["s"], dtype=np.int64)
    d = {}
    d["l"] = np.exp(rng.standard_normal(nl))
    nq = len(q)
    d["det"] = np.exp(rng.standard_normal(nq))
    q2 = [rng.standard_normal(k - 1) for k in q]
    d["q2"] = np.concatenate(q2) if q2 else np.zeros(0)
    d["q1"] = np.array([np.sqrt(d["det"][i] + q2[i] @ q2[i]) for i in range(nq)])
    d["auxdet"] = np.sqrt(2 * d["det"])
    d["auxtr"] = np.sqrt(2) * (d["q1"] + d["auxdet"])
    us, perms = [], []
    for n in s:
        n = int(n)
        if kind == "S2":
            Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
            ev = 10.0 ** rng.uniform(-6, 6, n)
            D = (Q * ev) @ Q.T
            D = (D + D.T) / 2
            U = np.linalg.cholesky(D).T
        else:
            R = np.linalg.qr(np.eye(n) + 0.3 * rng.standard_normal((n, n)), mode="r")
            U = np.triu(R) * np.sign(np.diag(R))[:, None]
        # d.u holds U in the upper triangle and mirrors it below (urotorder.c:400-401)
        full = U + np.triu(U, 1).T
        us.append(full.ravel(order="F"))
        perms.append(rng.permutation(n) + 1.0)
    d["u"] = np.concatenate(us) if us else np.zeros(0)
    d["perm"] = np.concatenate(perms).reshape(-1, 1) if perms else np.zeros((0, 0))
    return d


def synth_frames(s, seed: int = SEED0 + 7):
    """Spectral factor of a PSD iterate in SeDuMi's product form (what qrK leaves in vfrm.s,
    qrK.c:86-122): per block an n x n column-major array whose column k < n-1 holds the Householder
    vector q_k in rows k..n-1 and whose last column holds beta_0..beta_{n-2}, Q_k = I - q_k q_k'/beta_k
    (orthogonal iff beta_k = |q_k|^2/2); eigenvalue labels `lab` > 0.  Returns (lab, frms)."""
    rng = np.random.default_rng(seed)
    labs, frames = [], []
    for n in np.asarray(s, dtype=np.int64).ravel():
        n = int(n)
        F = np.zeros((n, n))
        for k in range(n - 1):
            q = rng.standard_normal(n - k)
            F[k:, k] = q
            F[k, n - 1] = 0.5 * float(q @ q)
        labs.append(np.exp(0.5 * rng.standard_normal(n)))
        frames.append(F.ravel(order="F"))
    if not labs:
        return np.zeros(0), np.zeros(0)
    return np.concatenate(labs), np.concatenate(frames)


def synth_hermitian_mixed(l=3, sreal=(4,), sherm=(3,), m=6, density=0.45, seed=SEED0 + 8):
    """A small problem directly in SeDuMi's INTERNAL form with Hermitian PSD blocks (K.s = [real..., Hermitian...],
    K.rsdpN = #real): At rows [x0, LP | vec(real blocks), lower triangle | per Hermitian block: Re part (lower
    triangle incl. diagonal), Im part (strictly lower)], as pretransfo.m:436-480 leaves them.
    Returns (At, b, c, K) with K already finished (cones.finish_K)."""
    from . import cones
    rng = np.random.default_rng(seed)
    K = cones.finish_K({"l": float(l), "q": np.zeros(0), "s": np.array(list(sreal) + list(sherm), dtype=float),
                        "rsdpN": len(sreal)})
    N = int(K["N"])
    rows, cols, vals = [], [], []
    for j in range(m):
        for r in range(1, l):
            if rng.random() < 0.7:
                rows.append(r); cols.append(j); vals.append(rng.standard_normal())
        off = l
        for n in sreal:
            for q in range(n):
                for p in range(q, n):
                    if rng.random() < density:
                        rows.append(off + p + q * n); cols.append(j); vals.append(rng.standard_normal())
            off += n * n
        for n in sherm:
            for q in range(n):
                for p in range(q, n):
                    if rng.random() < density:
                        rows.append(off + p + q * n); cols.append(j); vals.append(rng.standard_normal())
            for q in range(n):
                for p in range(q + 1, n):
                    if rng.random() < density:
                        rows.append(off + n * n + p + q * n); cols.append(j); vals.append(rng.standard_normal())
            off += 2 * n * n
    At = sp.csc_matrix((vals, (rows, cols)), shape=(N, m))
    At.sort_indices()
    return At, rng.standard_normal(m), rng.standard_normal(N), K


def scaling_hermitian(K: dict, seed: int = SEED0):
    """NT scaling for a cone with Hermitian blocks: like scaling(.., "S1") with complex upper-triangular factors
    (real positive diagonal) stored [vec Re U; vec Im U] for the blocks after K.rsdpN."""
    rng = np.random.default_rng(seed)
    s = np.asarray(K["s"], dtype=np.int64)
    nr = int(K.get("rsdpN", len(s)))
    d = {"l": np.exp(rng.standard_normal(int(K["l"]))), "det": np.zeros(0), "q1": np.zeros(0), "q2": np.zeros(0)}
    us, perms = [], []
    for i, n in enumerate(s):
        n = int(n)
        G = rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if i >= nr else 0)
        R = np.linalg.qr(np.eye(n) + 0.3 * G, mode="r")
        ph = np.diag(R) / np.abs(np.diag(R))
        U = np.triu(R) * np.conj(ph)[:, None]                 # positive real diagonal
        full = U + np.triu(U, 1).conj().T
        us.append(full.real.ravel(order="F"))
        if i >= nr:
            us.append(np.triu(U, 1).imag.ravel(order="F") - np.triu(U, 1).imag.T.ravel(order="F"))
        perms.append(rng.permutation(n) + 1.0)
    d["u"] = np.concatenate(us) if us else np.zeros(0)
    d["perm"] = np.concatenate(perms).reshape(-1, 1) if perms else np.zeros((0, 0))
    return d
