"""Supernodal symbolic Cholesky: ordering, elimination tree, column structures, supernodes.

Host-side (symbolic phase stays on the CPU, BASELINE.json north_star).  Produces the
same *data structure* the reference's symfctmex emits (symfctmex.c:100-126,192-254):

  L.perm   m x 1, 1-based doubles (fill-reducing order, etree-postordered)
  L.L      m x m sparse unit-lower pattern; per column: ascending rows, diagonal
           first; inside a supernode column j+1 has the pattern of column j minus
           its first row, so a supernode's values form a packed trapezoid
  L.xsuper (nsuper+1) x 1, 1-based doubles
  L.tmpsiz workspace bound for the reference's precorrect (choltmpsiz.c:57-101)

The ordering itself is not part of the parity contract (any valid triple is a valid
input to blkchol); we use SuperLU's MMD on A+A' via scipy, where the reference uses
Liu's MMD (ordmmd.c).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def mmd_order(S: sp.csc_matrix) -> np.ndarray:
    """Minimum-degree ordering of a symmetric pattern (0-based permutation)."""
    import scipy.sparse.linalg as spla
    m = S.shape[0]
    P = sp.csc_matrix((np.ones(S.nnz), S.indices, S.indptr), shape=S.shape)
    P = P + P.T
    P.setdiag(0)
    P.eliminate_zeros()
    deg = np.asarray(abs(P).sum(axis=1)).ravel()
    M = sp.csc_matrix(-P + sp.diags(deg + 1.0))
    lu = spla.splu(M, permc_spec="MMD_AT_PLUS_A", diag_pivot_thresh=0.0,
                   options=dict(SymmetricMode=True))
    return _perm_from_permc(lu.perm_c)


def _perm_from_permc(perm_c) -> np.ndarray:
    # SuperLU: column j of A is column perm_c[j] of A*Pc  ->  elimination order = argsort
    return np.argsort(np.asarray(perm_c, dtype=np.int64), kind="stable")


def etree(B: sp.csc_matrix) -> np.ndarray:
    """Elimination tree of a symmetric pattern (upper part used). parent[j] = -1 for roots."""
    m = B.shape[0]
    parent = np.full(m, -1, dtype=np.int64)
    anc = np.full(m, -1, dtype=np.int64)
    ip, ind = B.indptr, B.indices
    for j in range(m):
        for i in ind[ip[j]:ip[j + 1]]:
            i = int(i)
            while i != -1 and i < j:
                nxt = anc[i]
                anc[i] = j
                if nxt == -1:
                    parent[i] = j
                i = nxt
    return parent


def postorder(parent: np.ndarray) -> np.ndarray:
    m = len(parent)
    children = [[] for _ in range(m)]
    roots = []
    for j in range(m - 1, -1, -1):          # reversed so children pop in ascending order
        p = parent[j]
        (roots if p < 0 else children[p]).append(j)
    post = []
    for r in reversed(roots):
        stack = [(r, False)]
        while stack:
            v, done = stack.pop()
            if done:
                post.append(v)
            else:
                stack.append((v, True))
                for ch in children[v]:       # pushed descending -> popped ascending
                    stack.append((ch, False))
    return np.asarray(post, dtype=np.int64)


def column_structures(B: sp.csc_matrix, parent: np.ndarray):
    """Row structure (ascending, diagonal first) of every column of L."""
    m = B.shape[0]
    ip, ind = B.indptr, B.indices
    children = [[] for _ in range(m)]
    for j in range(m):
        if parent[j] >= 0:
            children[parent[j]].append(j)
    struct = [None] * m
    for j in range(m):
        col = ind[ip[j]:ip[j + 1]]
        parts = [col[col >= j], np.array([j])]
        for ch in children[j]:
            s = struct[ch]
            parts.append(s[1:])               # drop the child's diagonal
        struct[j] = np.unique(np.concatenate(parts))
    return struct


def tmpsiz_bound(ljc, lir, xsuper0) -> int:
    """choltmpsiz.c:57-101 (exact same quantity)."""
    nsuper = len(xsuper0) - 1
    m = xsuper0[-1]
    snode = np.zeros(m, dtype=np.int64)
    for s in range(nsuper):
        snode[xsuper0[s]:xsuper0[s + 1]] = s
    tmpsiz = 0
    for ksup in range(nsuper):
        k = xsuper0[ksup]
        inz = ljc[k] + (xsuper0[ksup + 1] - k)
        mk = ljc[k + 1] - inz
        ubsiz = mk * (mk + 1) // 2
        if mk == 0:
            continue
        i = lir[ljc[k + 1] - 1]
        while inz < ljc[k + 1] and ubsiz > tmpsiz:
            j = lir[inz]
            nextj = xsuper0[snode[j] + 1]
            if i < nextj:
                ncolup = mk
                inz = ljc[k + 1]
            else:
                ncolup = 1
                inz += 1
                while lir[inz] < nextj:
                    ncolup += 1
                    inz += 1
            tmpsiz = max(tmpsiz, mk * ncolup - ncolup * (ncolup - 1) // 2)
            mk -= ncolup
            ubsiz = mk * (mk + 1) // 2
    return int(tmpsiz)


def symbolic_factor(ADA: sp.csc_matrix, perm=None) -> dict:
    """(perm, L pattern, xsuper, tmpsiz) for the symmetric pattern ``ADA``.
    ``perm`` (0-based, optional) overrides the minimum-degree ordering."""
    m = ADA.shape[0]
    S = sp.csc_matrix(ADA)
    if perm is None:
        perm = mmd_order(S)
    perm = np.asarray(perm, dtype=np.int64).ravel()
    pat = sp.csc_matrix((np.ones(S.nnz), S.indices, S.indptr), shape=S.shape)
    pat = sp.csc_matrix(pat + pat.T)
    B = sp.csc_matrix(pat[perm][:, perm])
    B.sort_indices()
    par = etree(B)
    post = postorder(par)
    perm = perm[post]
    B = sp.csc_matrix(pat[perm][:, perm])
    B.sort_indices()
    par = etree(B)
    struct = column_structures(B, par)
    cc = np.array([len(s) for s in struct], dtype=np.int64)
    xsuper = [0]
    for j in range(1, m):
        same = par[j - 1] == j and cc[j] == cc[j - 1] - 1      # maximal supernodes
        if not same:
            xsuper.append(j)
    xsuper.append(m)
    xsuper = np.asarray(xsuper, dtype=np.int64)
    ljc = np.r_[0, np.cumsum(cc)]
    lir = np.concatenate(struct) if m else np.zeros(0, np.int64)
    L = sp.csc_matrix((np.ones(lir.size), lir, ljc), shape=(m, m))
    return {"perm": (perm + 1).astype(np.float64).reshape(-1, 1),
            "L": L,
            "xsuper": (xsuper + 1).astype(np.float64).reshape(-1, 1),
            "tmpsiz": float(tmpsiz_bound(ljc, lir, xsuper))}


# --------------------------------------------------------------------------- dense columns (symbcholden.m:45-62)
def symbfwblk(L: dict, B: sp.csc_matrix) -> sp.csc_matrix:
    """Pattern of L.L \\ B(L.perm,:) for a sparse B (symbfwblk.c): the nonzeros of every column spread along the
    elimination tree -- row i reaches its parent, the first off-diagonal row of column i of L.L, and so on."""
    LL = sp.csc_matrix(L["L"])
    m = LL.shape[0]
    perm = np.asarray(L["perm"], dtype=np.int64).ravel() - 1
    invperm = np.empty(m, dtype=np.int64)
    invperm[perm] = np.arange(m)
    ip, ind = LL.indptr, LL.indices
    parent = np.full(m, -1, dtype=np.int64)
    for j in range(m):
        if ip[j + 1] - ip[j] > 1:
            parent[j] = ind[ip[j] + 1]
    B = sp.csc_matrix(B)
    cols, jc = [], [0]
    for c in range(B.shape[1]):
        mark = np.zeros(m, dtype=bool)
        for r in B.indices[B.indptr[c]:B.indptr[c + 1]]:
            j = int(invperm[r])
            while j != -1 and not mark[j]:
                mark[j] = True
                j = int(parent[j])
        rows = np.flatnonzero(mark)
        cols.append(rows)
        jc.append(jc[-1] + rows.size)
    ir = np.concatenate(cols) if cols else np.zeros(0, dtype=np.int64)
    return sp.csc_matrix((np.ones(ir.size), ir, np.asarray(jc)), shape=(m, B.shape[1]))


def symbcholden(L: dict, dense, DAt_denq=None) -> dict:
    """Symbolic product-form structure of the dense columns (symbcholden.m:45-62 with finsymbden.c:57-84,155-175) for
    LP dense columns (no dense Lorentz blocks): Lden = {LAD (pattern of L\\Ad), perm, dz, first}."""
    from . import setup as hsetup
    assert len(dense.q) == 0, "dense Lorentz blocks: use the reference's symbcholden"
    nden = len(dense.cols)
    LAD = symbfwblk(L, sp.csc_matrix(dense.A[:, :nden]))
    perm, dz = hsetup.incorder(LAD)                       # incremental ordering of the columns (incorder.c)
    m = LAD.shape[0]
    dz = sp.csc_matrix(dz)
    # first affecting pivot of each column (getfirstpiv): position of its earliest row in the dz order, then the
    # first dz column whose cumulative count covers it
    invdz = np.zeros(m, dtype=np.int64)
    invdz[dz.indices] = np.arange(dz.nnz)
    first = np.empty(nden)
    for j in range(nden):
        rows = LAD.indices[LAD.indptr[j]:LAD.indptr[j + 1]]
        if rows.size == 0:
            first[j] = nden + 1.0
            continue
        fj = int(invdz[rows].min())
        y = int(np.searchsorted(dz.indptr[1:], fj + 1, side="left"))
        first[j] = y + 1.0
    return {"LAD": LAD, "perm": np.asarray(perm, dtype=np.float64).reshape(-1, 1), "dz": dz, "first": first.reshape(-1, 1)}
