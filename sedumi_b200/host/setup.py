"""Once-per-solve setup that produces every structure the hot path consumes.

Host-side restatement (numpy/scipy) of sedumi.m:356-407 and of the small symbolic
helpers it calls, so the hot path can be driven here without MATLAB/Octave:

  partitA.c:73-146   -> :func:`partitA`
  extractA.c         -> :func:`extractA`
  findblks.c         -> :func:`findblks`
  sortnnz.c:60-70    -> :func:`sortnnz`   (stable; the reference's qsort tie order is unspecified)
  incorder.c:140-200 -> :func:`incorder`
  getdense.m:46-100  -> :func:`getdense`
  getsymbada.m:41-60 -> :func:`getsymbada`
  symbchol.m:57-83   -> :func:`symbchol`   (dense shortcut or :mod:`.symbolic`)
  sdinit.m:42-104    -> :func:`sdinit_scaling` (the scaling part only)

All outputs use the reference's conventions (1-based doubles for perm/xsuper/blkstart,
0-based doubles for Ablkjc), see SURVEY.md Appendix A.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from . import symbolic
from .cones import K_for_mex


def spars(X) -> float:
    m, n = X.shape
    return X.nnz / max(m * n, 1)


def partitA(At: sp.csc_matrix, blkstart) -> np.ndarray:
    """Ablkjc (m x len(blkstart)), 0-based absolute offsets into At.indices: first
    nonzero of column j with row >= blkstart[k]-1 (partitA.c:73-84,137-146)."""
    blk0 = np.asarray(blkstart, dtype=np.int64).ravel() - 1
    m = At.shape[1]
    out = np.zeros((m, len(blk0)), dtype=np.float64)
    ip, ind = At.indptr, At.indices
    for j in range(m):
        col = ind[ip[j]:ip[j + 1]]
        out[j, :] = ip[j] + np.searchsorted(col, blk0, side="left")
    return out


def _col_ranges(At, Ablkjc, blk0, blk1):
    m = At.shape[1]
    lo = At.indptr[:-1].astype(np.int64) if blk0 < 1 else Ablkjc[:, blk0 - 1].astype(np.int64)
    hi = At.indptr[1:].astype(np.int64) if (blk1 is None or blk1 > Ablkjc.shape[1]) \
        else Ablkjc[:, blk1 - 1].astype(np.int64)
    return lo, hi


def extractA(At, Ablkjc, blk0, blk1, row0, row1=None) -> sp.csc_matrix:
    """At(row0:row1-1, :) restricted to the per-column nz range (extractA.m:2-12).
    row0/row1 are 1-based like MATLAB."""
    if row1 is None:
        row1 = At.shape[0] + 1
    lo, hi = _col_ranges(At, Ablkjc, blk0, blk1)
    m = At.shape[1]
    cnt = hi - lo
    indptr = np.r_[0, np.cumsum(cnt)]
    sel = np.concatenate([np.arange(lo[j], hi[j]) for j in range(m)]) if cnt.sum() else np.zeros(0, np.int64)
    return sp.csc_matrix((At.data[sel], At.indices[sel] - (row0 - 1), indptr),
                         shape=(int(row1 - row0), m))


def findblks(At, Ablkjc, blk0, blk1, blkstart) -> sp.csc_matrix:
    """Block-incidence matrix (nblk x m) of the rows in the given nz range, partitioned
    by ``blkstart`` (1-based, length nblk+1) (findblks.m:2-8)."""
    bs = np.asarray(blkstart, dtype=np.int64).ravel() - 1
    nblk = max(len(bs) - 1, 0)
    m = At.shape[1]
    lo, hi = _col_ranges(At, Ablkjc, blk0, blk1)
    rows, cols = [], []
    for j in range(m):
        r = At.indices[lo[j]:hi[j]]
        if r.size and nblk:
            k = np.unique(np.searchsorted(bs, r, side="right") - 1)
            k = k[(k >= 0) & (k < nblk)]
            rows.append(k)
            cols.append(np.full(k.size, j))
    if rows:
        rows, cols = np.concatenate(rows), np.concatenate(cols)
    else:
        rows, cols = np.zeros(0, np.int64), np.zeros(0, np.int64)
    return sp.csc_matrix((np.ones(rows.size), (rows, cols)), shape=(nblk, m))


def sortnnz(At, Ajc1, Ajc2) -> np.ndarray:
    """1-based permutation sorting columns by nnz in [Ajc1, Ajc2) (sortnnz.c:60-70)."""
    lo = At.indptr[:-1] if Ajc1 is None or np.size(Ajc1) == 0 else np.asarray(Ajc1).ravel()
    hi = At.indptr[1:] if Ajc2 is None or np.size(Ajc2) == 0 else np.asarray(Ajc2).ravel()
    cnt = np.asarray(hi, dtype=np.int64) - np.asarray(lo, dtype=np.int64)
    return (np.argsort(cnt, kind="stable") + 1).astype(np.float64)


def incorder(At, Ajc1=None, ifirst=1):
    """Greedy sparsest-first ordering of the PSD part and the incremental pattern dz
    (incorder.c:140-200).  Returns (perm 1-based float (m,), dz CSC N x m pattern)."""
    N, m = At.shape
    first = int(ifirst) - 1
    ip, ind = At.indptr, At.indices
    lo = ip[:-1].astype(np.int64) if Ajc1 is None else np.asarray(Ajc1, dtype=np.int64).ravel()
    hi = ip[1:].astype(np.int64)
    lenud = N - first
    # transpose of the PSD part: for each subscript, the constraints that use it
    rows = np.concatenate([ind[lo[j]:hi[j]] for j in range(m)]) if m else np.zeros(0, np.int64)
    cols = np.repeat(np.arange(m), hi - lo)
    T = sp.csr_matrix((np.ones(rows.size, dtype=np.int8), (rows - first, cols)), shape=(max(lenud, 0), m))
    remaining = (hi - lo).astype(np.int64)
    perm = np.arange(m)
    discard = np.zeros(max(lenud, 0), dtype=bool)
    dzjc = np.zeros(m + 1, dtype=np.int64)
    dzir = []
    for k in range(m):
        cand = perm[k:]
        kmin = k + int(np.argmin(remaining[cand]))        # first minimum, like the C loop
        perm[k], perm[kmin] = perm[kmin], perm[k]
        pk = perm[k]
        sub = ind[lo[pk]:hi[pk]] - first
        new = sub[~discard[sub]]
        discard[new] = True
        dzir.append(new + first)
        dzjc[k + 1] = dzjc[k] + new.size
        if new.size:
            # every constraint using a newly covered subscript gets shorter
            hit = T[new].indices
            np.subtract.at(remaining, hit, 1)
    dzir = np.concatenate(dzir) if dzir else np.zeros(0, np.int64)
    dz = sp.csc_matrix((np.ones(dzir.size), dzir, dzjc), shape=(N, m))
    return (perm + 1).astype(np.float64), dz


@dataclass
class Dense:
    cols: np.ndarray = field(default_factory=lambda: np.zeros(0))   # 1-based row ids of At
    q: np.ndarray = field(default_factory=lambda: np.zeros(0))      # 1-based Lorentz block ids
    l: int = 0
    A: sp.csc_matrix | None = None                                  # m x len(cols)

    def for_mex(self):
        return {"cols": self.cols.reshape(-1, 1).astype(np.float64),
                "q": self.q.reshape(-1, 1).astype(np.float64),
                "l": float(self.l), "A": self.A}


def getdense(At, Ablkjc, K, denq=0.75, denf=10.0):
    """Dense-column detection (getdense.m:46-100)."""
    NORMDEN = 5
    N, m = At.shape
    lq = int(K["lq"])
    Alq = extractA(At, Ablkjc, 0, 3, 1, lq + 1)
    colnz = np.asarray((Alq != 0).sum(axis=1)).ravel().astype(np.float64)
    blk = findblks(At, Ablkjc, 3, None, K["sblkstart"])
    h = max([NORMDEN] + list(np.asarray(blk.sum(axis=1)).ravel()))
    i1, i2 = int(K["mainblks"][0]), int(K["mainblks"][1])
    Ablkq = extractA(At, Ablkjc, 1, 2, i1, i2)
    if i1 < i2:
        Ablkq2 = findblks(At, Ablkjc, 2, 3, K["qblkstart"])
        Ablkq = sp.csc_matrix(((Ablkq != 0).astype(float) + Ablkq2) != 0, dtype=np.float64)
        colnz[i1 - 1:i2 - 1] = np.asarray(Ablkq.sum(axis=1)).ravel()
    big = colnz[colnz > h]
    denqN = int(np.ceil(denq * len(colnz))) - (N - len(big))
    spquant = h if denqN < 1 else np.sort(big)[denqN - 1]
    cols = np.flatnonzero(colnz > denf * spquant) + 1
    dq = np.flatnonzero(colnz[i1 - 1:i2 - 1] > denf * spquant) + 1
    dl = int((cols < i1).sum())
    if len(cols) > m / 2:
        cols, dq, dl = np.zeros(0, np.int64), np.zeros(0, np.int64), 0
    if len(dq) == 0:
        Adotdden = sp.csc_matrix((m, 0))
    else:
        Adotdden = sp.csc_matrix(Ablkq[dq - 1, :].T)
    return Dense(cols=cols.astype(np.float64), q=dq.astype(np.float64), l=dl), Adotdden


def getsymbada(At, Ablkjc, DAtq, psdblkstart) -> sp.csc_matrix:
    """Pattern of ADA; all-ones when >90 % dense (getsymbada.m:41-60)."""
    m = At.shape[1]
    ones = lambda: sp.csc_matrix(np.ones((m, m)))
    Alpq = sp.csc_matrix(extractA(At, Ablkjc, 0, 3, 1, int(psdblkstart[0])) != 0, dtype=np.float64)
    Ablks = findblks(At, Ablkjc, 3, None, psdblkstart)
    if (Ablks.shape[0] and spars(Ablks) == 1) or (Alpq.shape[0] and spars(Alpq) == 1) or \
            (DAtq is not None and DAtq.shape[0] and spars(DAtq) == 1):
        return ones()
    S = sp.csc_matrix((m, m))
    if DAtq is not None and DAtq.shape[0]:
        P = sp.csc_matrix(DAtq != 0, dtype=np.float64)
        S = sp.csc_matrix(P.T @ P)
        if spars(S) > 0.9:
            return ones()
    S = sp.csc_matrix(S + Alpq.T @ Alpq)
    if spars(S) > 0.9:
        return ones()
    S = sp.csc_matrix(S + Ablks.T @ Ablks)
    S.data[:] = 1.0
    S.sort_indices()
    return S


def symbchol(ADA: sp.csc_matrix, perm=None) -> dict:
    """Symbolic Cholesky structure L.{perm, L, xsuper, tmpsiz} (symbchol.m:57-83).

    Fully dense pattern -> the reference's one-supernode shortcut (symbchol.m:75-77).
    Otherwise a fill-reducing ordering + supernodal symbolic factorisation from
    :mod:`.symbolic` (the reference uses ordmmd.c/symfct.c here; any valid
    (perm, L, xsuper) triple is an equally valid input for the numeric phase)."""
    m = ADA.shape[0]
    if spars(ADA) >= 1:
        return {"perm": np.arange(1, m + 1, dtype=np.float64).reshape(-1, 1),
                "L": sp.csc_matrix(np.tril(np.ones((m, m)))),
                "xsuper": np.array([1.0, m + 1.0]).reshape(-1, 1),
                "tmpsiz": 0.0}
    return symbolic.symbolic_factor(ADA, perm)


def L_for_mex(L: dict) -> dict:
    keys = ["perm", "L", "xsuper", "tmpsiz", "d", "skip", "add"]
    return {k: L[k] for k in keys if k in L}


def sdinit_scaling(b, c, K, mu_par=1.0) -> dict:
    """Iteration-0 NT scaling ``d`` (sdinit.m:46-78)."""
    maxb = np.abs(b).max() if b.size else 0.0
    maxc = np.abs(c).max() if c.size else 0.0
    mu = mu_par * np.sqrt((1 + maxb) * (1 + maxc))
    d0 = np.sqrt((1 + maxb) / (1 + maxc))
    x0 = mu_par
    z0 = mu ** 2 / x0
    nq = len(K["q"])
    d = {}
    d["l"] = d0 ** 2 * np.ones(int(K["l"]))
    d["l"][0] = x0 / z0
    d["det"] = d0 ** 2 * np.ones(nq)
    d["q1"] = np.sqrt(2) * d0 * np.ones(nq)
    d["q2"] = np.zeros(int(K["mainblks"][2] - K["mainblks"][1]))
    d["auxdet"] = np.sqrt(2 * d["det"])
    d["auxtr"] = np.sqrt(2) * (d["q1"] + d["auxdet"])
    u = [np.sqrt(d0) * np.eye(int(n)).ravel(order="F") for n in K["s"]]
    d["u"] = np.concatenate(u) if u else np.zeros(0)
    d["perm"] = np.zeros((0, 0))
    return d


@dataclass
class HotPathSetup:
    """Everything iteration-invariant that the hot path consumes (sedumi.m:356-392)."""
    At: sp.csc_matrix
    b: np.ndarray
    c: np.ndarray
    K: dict
    Ablkjc: np.ndarray
    dense: Dense
    DAt_denq: sp.csc_matrix
    Aord: dict
    ADA: sp.csc_matrix           # pattern (values all one)
    L: dict

    @property
    def m(self):
        return self.At.shape[1]

    def Kmex(self):
        return K_for_mex(self.K)


def build_setup(At, b, c, K, denq=0.75, denf=10.0, perm=None) -> HotPathSetup:
    """sedumi.m:356-392 on an internal-form problem (output of cones.pretransfo)."""
    At = sp.csc_matrix(At)
    At.sort_indices()
    m = At.shape[1]
    Ablkjc = partitA(At, K["mainblks"])
    dense, denq_pat = getdense(At, Ablkjc, K, denq, denf)
    if len(dense.cols):
        rows = dense.cols.astype(np.int64) - 1
        dense.A = sp.csc_matrix(At[rows, :].T)
        At = sp.lil_matrix(At)
        At[rows, :] = 0.0
        At = sp.csc_matrix(At)
        At.eliminate_zeros()
        At.sort_indices()
        Ablkjc = partitA(At, K["mainblks"])
    else:
        dense.A = sp.csc_matrix((m, 0))
    Aord = {"lqperm": sortnnz(At, None, Ablkjc[:, 2]).reshape(-1, 1)}
    DAtq = findblks(At, Ablkjc, 2, 3, K["qblkstart"])
    if DAtq.shape[0]:
        if len(dense.q):
            DAtq = sp.lil_matrix(DAtq)
            DAtq[dense.q.astype(np.int64) - 1, :] = 0.0
            DAtq = sp.csc_matrix(DAtq)
            DAtq.eliminate_zeros()
        tr = extractA(At, Ablkjc, 1, 2, int(K["mainblks"][0]), int(K["mainblks"][1]))
        DAtq = sp.csc_matrix(DAtq + sp.csc_matrix(tr != 0, dtype=np.float64))
        Aord["qperm"] = sortnnz(DAtq, None, None).reshape(-1, 1)
    else:
        Aord["qperm"] = np.arange(1, m + 1, dtype=np.float64).reshape(-1, 1)
    sperm, dz = incorder(At, Ablkjc[:, 2], int(K["mainblks"][2]))
    Aord["sperm"] = sperm.reshape(-1, 1)
    Aord["dz"] = dz
    ADA = getsymbada(At, Ablkjc, DAtq, K["sblkstart"])
    L = symbchol(ADA, perm)
    return HotPathSetup(At=At, b=b, c=c, K=K, Ablkjc=Ablkjc, dense=dense, DAt_denq=denq_pat,
                        Aord=Aord, ADA=ADA, L=L)
