"""CPU restatement ("port") of the M-code pieces of SeDuMi's hot path -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product path (sedumi_b200/) never does.

The C pieces of the path are NOT restated: the oracle for those is the reference's own C,
compiled unmodified into oracle/_ref by oracle/Makefile.  What is restated here (numpy/scipy)
is the MATLAB glue that has no C source and cannot run without MATLAB/Octave:

  psdscale.m:45-119   -> psdscale        getada.m:14-40     -> getada_m
  getDAtm.m:39-47     -> getDAtm         deninfac.m:57-93   -> deninfac
  Amul.m:42-56        -> Amul            asmDxq.m:40-68     -> asmDxq
  wrapPcg.m:42-94     -> wrapPcg_onepass (the part before PCG refinement)

Parity status: these follow the M source line by line; MATLAB itself is not available here,
so they are pinned only through the reference MEX functions they call (oracle/_ref) and
through algebraic identities checked in tests/test_oracle.py -- "parity unpinned" in the
sense of the task statement for the pure-M arithmetic (dense products, sparse products).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def psdscale(ud, x, K, transp=False):
    """y = psdscale(ud,x,K,transp)  (psdscale.m:45-119): per block Y = T'*X*T with T = tril/triu(U); the
    blocks after K.rsdpN are Hermitian, stored [vec Re; vec Im], and get the imaginary diagonal of Y zeroed
    (psdscale.m:111-116)."""
    Ks = np.asarray(K["s"], dtype=np.int64).ravel()
    if Ks.size == 0:
        return np.zeros(0)
    nr = int(K.get("rsdpN", Ks.size)) if isinstance(K, dict) else Ks.size
    perm = None
    if isinstance(ud, dict):
        p = np.asarray(ud.get("perm", np.zeros(0))).ravel()
        perm = p.astype(np.int64) - 1 if p.size else None
        u = np.asarray(ud["u"], dtype=float).ravel()
    else:
        u = np.asarray(ud, dtype=float).ravel()
    x = np.asarray(x, dtype=float).ravel()
    N = int((Ks ** 2).sum() + (Ks[nr:] ** 2).sum())
    xi = x.size - N
    y = np.zeros(N)
    ui = yi = pi = 0
    for i, n in enumerate(Ks):
        n = int(n)
        q = n * n
        cplx = i >= nr
        TT = u[ui:ui + q].reshape(n, n, order="F")
        ui += q
        if cplx:
            TT = TT + 1j * u[ui:ui + q].reshape(n, n, order="F")
            ui += q
        TT = np.triu(TT) if transp else np.tril(TT)
        XX = x[xi:xi + q].reshape(n, n, order="F")
        xi += q
        if cplx:
            XX = XX + 1j * x[xi:xi + q].reshape(n, n, order="F")
            xi += q
        if perm is not None and not transp:
            PP = perm[pi:pi + n]
            pi += n
            XX = XX[np.ix_(PP, PP)]
        XX = TT.conj().T @ XX @ TT
        if perm is not None and transp:
            PP = perm[pi:pi + n]
            pi += n
            Z = np.empty_like(XX)
            Z[np.ix_(PP, PP)] = XX
            XX = Z
        y[yi:yi + q] = XX.real.ravel(order="F")
        yi += q
        if cplx:
            Z = XX.imag.copy()
            Z[np.diag_indices(n)] = 0.0
            y[yi:yi + q] = Z.ravel(order="F")
            yi += q
    return y


def invcholfac_dense(u, K, perm=None):
    """Independent dense formula for invcholfac (used to cross-check oracle/_ref/invcholfac.so)."""
    Ks = np.asarray(K["s"], dtype=np.int64).ravel()
    u = np.asarray(u, dtype=float).ravel()
    out, ui, pi = [], 0, 0
    perm = None if perm is None or np.size(perm) == 0 else np.asarray(perm).ravel().astype(np.int64) - 1
    for n in Ks:
        n = int(n)
        U = np.triu(u[ui:ui + n * n].reshape(n, n, order="F"))
        ui += n * n
        Z = U.T @ U
        if perm is not None:
            PP = perm[pi:pi + n]
            pi += n
            Y = np.empty_like(Z)
            Y[np.ix_(PP, PP)] = Z
            Z = Y
        out.append(Z.ravel(order="F"))
    return np.concatenate(out) if out else np.zeros(0)


def ada_dense_formula(At, K, d, udsqr):
    """Independent dense evaluation of ADA = A*D(d^2)*A' (all cones) for cross-checks:
    LP: d.l ; Lorentz: det*(J) + (Dq a)(Dq a)' handled via getada1/2 definitions ; PSD: <A_i, D A_j D>."""
    At = sp.csc_matrix(At)
    N, m = At.shape
    A = At.toarray()
    lpN = int(K["l"])
    q = np.asarray(K["q"], dtype=np.int64)
    s = np.asarray(K["s"], dtype=np.int64)
    nq = len(q)
    ADA = (A[:lpN].T * d["l"]) @ A[:lpN]
    r = lpN + nq
    for k in range(nq):
        nk = int(q[k])
        x1 = A[lpN + k]                        # trace row
        x2 = A[r:r + nk - 1]
        det = d["det"][k]
        ADA += det * (x2.T @ x2 - np.outer(x1, x1))
        dq = d["q1"][k] * x1 + d["q2"][r - lpN - nq: r - lpN - nq + nk - 1] @ x2
        ADA += np.outer(dq, dq)
        r += nk - 1
    off = 0
    for n in s:
        n = int(n)
        D = np.asarray(udsqr[off:off + n * n]).reshape(n, n, order="F")
        V = A[r:r + n * n]                     # folded lower-triangular coefficients
        W = np.empty_like(V)
        for j in range(m):
            X = V[:, j].reshape(n, n, order="F")
            X = (X + X.T) / 2
            W[:, j] = (D @ X @ D).ravel(order="F")
        ADA += V.T @ W
        r += n * n
        off += n * n
    return ADA


def getada_m(At, K, d, DAtq, pattern=None):
    """getada.m:14-40 (the M path sedumi.m:446-448 takes when sum(K.s)==0):
        ADA = DAt.q'*DAt.q + Alq'*diag([d.l; -d.det; d.det(k) on the norm-bound rows of cone k])*Alq,
        absd = diag(ADA).
    Returns (ADA as CSC on `pattern` if given, absd)."""
    import scipy.sparse as sp
    nl = int(K["l"])
    q = np.asarray(K["q"], dtype=np.int64)
    nq = len(q)
    lq = nl + int(q.sum())
    sv = np.r_[np.asarray(d["l"], dtype=np.float64).ravel(), -np.asarray(d["det"], dtype=np.float64).ravel(),
               np.repeat(np.asarray(d["det"], dtype=np.float64).ravel(), q - 1) if nq else np.zeros(0)]
    Alq = sp.csc_matrix(At)[:lq, :]
    ADA = (DAtq.T @ DAtq + Alq.T @ sp.diags(sv) @ Alq).toarray()
    absd = np.diag(ADA).copy()
    if pattern is None:
        return sp.csc_matrix(ADA), absd
    P = sp.csc_matrix(pattern)
    rows = P.indices
    cols = np.repeat(np.arange(P.shape[1]), np.diff(P.indptr))
    return sp.csc_matrix((ADA[rows, cols], P.indices.copy(), P.indptr.copy()), shape=P.shape), absd


# --------------------------------------------------------------------------- PSD algebra that is M code in the reference
def _blocks(x, K):
    Ks = np.asarray(K["s"], dtype=np.int64).ravel()
    N = int((Ks ** 2).sum())
    x = np.asarray(x, dtype=float).ravel()
    off = x.size - N
    out = []
    for n in Ks:
        n = int(n)
        out.append(x[off:off + n * n].reshape(n, n, order="F"))
        off += n * n
    return out


def psdjmul(x, y, K):
    """psdjmul.m:38-74 (real blocks): Z_k = (X_k Y_k + (X_k Y_k)')/2."""
    return np.concatenate([(0.5 * (X @ Y + (X @ Y).T)).ravel(order="F") for X, Y in zip(_blocks(x, K), _blocks(y, K))])


def triumtriu(x, y, K):
    """triumtriu.m:38-73 (real blocks): Z_k = triu(X_k) triu(Y_k), upper triangle mirrored below."""
    out = []
    for X, Y in zip(_blocks(x, K), _blocks(y, K)):
        Z = np.triu(X) @ np.triu(Y)
        out.append((Z + np.triu(Z, 1).T).ravel(order="F"))
    return np.concatenate(out)


def psdfactor(x, K):
    """psdfactor.m:37-82 (real blocks): (ux, ispos); ux = L + tril(L,-1)' per block, zeros from the first block that is
    not positive definite on."""
    blocks = _blocks(x, K)
    out = [np.zeros(B.size) for B in blocks]
    for i, X in enumerate(blocks):
        try:
            L = np.linalg.cholesky(np.tril(X) + np.tril(X, -1).T)      # chol(.,'lower') reads the lower triangle
        except np.linalg.LinAlgError:
            return np.concatenate(out), False
        out[i] = (L + np.tril(L, -1).T).ravel(order="F")
    return np.concatenate(out), True


def psdinvscale(ud, x, K):
    """psdinvscale.m:37-83 (real blocks): Y_k = T \\ (X_k / T'), T = triu(U_k)."""
    import scipy.linalg as sla
    out = []
    for U, X in zip(_blocks(ud, K), _blocks(x, K)):
        T = np.triu(U)
        W = sla.solve_triangular(T, X.T, lower=False).T                 # X / T'
        out.append(sla.solve_triangular(T, W, lower=False).ravel(order="F"))
    return np.concatenate(out)


def psdeig(x, K, vectors=False):
    """psdeig.m:40-96 (real blocks): lab = 0.5 eig(XX + XX') per block, ascending; optionally the eigenvectors.
    The reference calls the host's eig(): LAPACK here (numpy.linalg.eigh) -- parity unpinned beyond that."""
    labs, qs = [], []
    for X in _blocks(x, K):
        w, Q = np.linalg.eigh(X + X.T)
        labs.append(0.5 * w)
        qs.append(Q.ravel(order="F"))
    lab = np.concatenate(labs) if labs else np.zeros(0)
    return (lab, np.concatenate(qs)) if vectors else lab


def minpsdeig(x, K):
    """minpsdeig.m:43-68: the smallest spectral coefficient over all blocks."""
    return float(psdeig(x, K).min())

