"""Cone bookkeeping and the real-data subset of SeDuMi's ``pretransfo``.

Host-side logic (stays on the CPU by design): it turns a
user problem ``(At, b, c, K)`` into SeDuMi's *internal* layout, which is what every
hot-path plugin consumes:

    x = [x0 ; LP ; Lorentz traces ; Lorentz norm parts ; vec(PSD blocks)]

Mirrors pretransfo.m:44-545 for real data (K.f, K.l, K.q, K.r, real K.s).  Complex
data (K.xcomplex / K.scomplex / K.ycomplex / K.z) is a "next" row
and raises NotImplementedError here.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _row(v) -> np.ndarray:
    return np.atleast_1d(np.asarray(v, dtype=np.float64)).ravel()


def normalise_K(K: dict) -> dict:
    """Fill in missing cone fields (pretransfo.m:44-92)."""
    out = {}
    out["f"] = int(_row(K.get("f", 0))[0]) if np.size(K.get("f", 0)) else 0
    out["l"] = int(_row(K.get("l", 0))[0]) if np.size(K.get("l", 0)) else 0
    for name, lo in (("q", 2), ("r", 3), ("s", 1)):
        v = _row(K.get(name, []))
        v = v[v != 0] if not np.count_nonzero(v) else v
        if not np.count_nonzero(v):
            v = np.zeros(0)
        if np.any(v != np.floor(v)) or np.any(v < lo):
            raise ValueError(f"K.{name} should contain only integers >= {lo}")
        out[name] = v.astype(np.int64)
    for name in ("xcomplex", "scomplex", "ycomplex", "z"):
        if name in K and np.size(K[name]) and np.count_nonzero(_row(K[name])):
            raise NotImplementedError(f"K.{name}: complex data is outside round-1 scope")
    return out


def pretransfo(At, b, c, K: dict, free: int | None = None, sdp: bool = True):
    """Internal-form problem ``(At, b, c, K)`` as produced by pretransfo.m.

    Returns ``At`` (N x m CSC, PSD coefficients folded into the lower triangle,
    pretransfo.m:434-454), dense ``b`` (m,), dense ``c`` (N,), and ``K`` with the
    detailed fields of pretransfo.m:486-542 (all index fields 1-based like MATLAB).
    """
    K = normalise_K(K)
    Kf, Kl, Kq, Kr, Ks = K["f"], K["l"], K["q"], K["r"], K["s"]
    N_fl = Kf + Kl
    L_q, N_q = len(Kq), int(Kq.sum())
    L_r, N_r = len(Kr), int(Kr.sum())
    L_qr = L_q + L_r
    L_s, N_s = len(Ks), int((Ks ** 2).sum())
    N_flqr = N_fl + N_q + N_r
    N = N_flqr + N_s

    At = sp.csc_matrix(At, dtype=np.float64)
    if At.shape[0] != N:
        if At.shape[1] == N:
            At = sp.csc_matrix(At.T)
        else:
            raise ValueError("(At,K) size mismatch")
    m = At.shape[1]
    b = np.asarray(b.todense() if sp.issparse(b) else b, dtype=np.float64).ravel()
    c = np.asarray(c.todense() if sp.issparse(c) else c, dtype=np.float64).ravel()
    if b.size != m:
        raise ValueError("(At,b) size mismatch")
    if c.size != N:
        raise ValueError("(c,K) size mismatch")

    # ---- diagonal PSD blocks become LP variables (pretransfo.m:231-246)
    if L_s and sdp:
        strt = np.cumsum(np.r_[0, Ks[:-1] ** 2])                 # 0-based block starts
        rows_nz = np.zeros(N_s, dtype=bool)
        rows_nz[np.unique(At.indices[At.indices >= N_flqr]) - N_flqr] = True
        rows_nz |= c[N_flqr:] != 0
        spattern = np.flatnonzero(rows_nz)
        blk = np.searchsorted(strt, spattern, side="right") - 1
        offdiag = (spattern - strt[blk]) % (Ks[blk] + 1) != 0
        sdiag = np.ones(L_s, dtype=bool)
        sdiag[blk[offdiag]] = False
    else:
        sdiag = Ks == 1
    sreal = ~sdiag

    ii, jj, vv = [], [], []            # 0-based (row in new x, col in old x)
    newL = 0
    newQ = np.zeros(0, dtype=np.int64)
    L_qrsz = L_qr + L_s
    if free is None or (free == 2 and L_qrsz):
        free = 1
    if Kf and not free:                # split free vars (pretransfo.m:337-345)
        jt = np.repeat(np.arange(Kf), 2)
        vt = np.tile([1.0, -1.0], Kf)
        ii.append(np.arange(2 * Kf)); jj.append(jt); vv.append(vt)
        newL = 2 * Kf
    if Kl:
        ii.append(newL + np.arange(Kl)); jj.append(Kf + np.arange(Kl)); vv.append(np.ones(Kl))
        newL += Kl
    if sdiag.any():                    # pretransfo.m:356-370
        jstrt_all = N_flqr + np.cumsum(np.r_[0, Ks[:-1] ** 2])
        for k in np.flatnonzero(sdiag):
            n = int(Ks[k])
            ii.append(newL + np.arange(n))
            jj.append(jstrt_all[k] + (n + 1) * np.arange(n))
            vv.append(np.ones(n))
            newL += n
    tr_off = newL
    nb_off = newL + L_qr
    if Kf and free:                    # free vars into a Lorentz cone (pretransfo.m:375-384)
        tr_off += 1
        nb_off += 1
        ii.append(nb_off + np.arange(Kf)); jj.append(np.arange(Kf)); vv.append(np.ones(Kf))
        nb_off += Kf
        newQ = np.array([Kf + 1], dtype=np.int64)
    if N_q:                            # trace block + norm-bound blocks (pretransfo.m:387-403)
        ndxs = np.cumsum(np.r_[0, Kq[:-1]])
        it = np.full(N_q, -1, dtype=np.int64)
        it[ndxs] = tr_off + np.arange(L_q)
        it[it < 0] = nb_off + np.arange(N_q - L_q)
        ii.append(it); jj.append(Kf + Kl + np.arange(N_q)); vv.append(np.ones(N_q))
        tr_off += L_q
        nb_off += N_q - L_q
    if N_r:                            # rotated cones (pretransfo.m:407-429)
        col0 = Kf + Kl + N_q
        nbpos = nb_off
        for k in range(L_r):
            n = int(Kr[k])
            base = col0 + int(Kr[:k].sum())
            tr = tr_off + k
            s = np.sqrt(0.5)
            # new trace = (x1+x2)/sqrt2 ; first norm-bound entry = (x1-x2)/sqrt2
            ii.append(np.array([tr, tr, nbpos, nbpos]))
            jj.append(np.array([base, base + 1, base, base + 1]))
            vv.append(np.array([s, s, s, -s]))
            ii.append(nbpos + 1 + np.arange(n - 2)); jj.append(base + 2 + np.arange(n - 2))
            vv.append(np.ones(n - 2))
            nbpos += n - 1
        nb_off += N_r - L_r
    if sreal.any():                    # fold to lower triangle (pretransfo.m:434-454)
        jstrt_all = N_flqr + np.cumsum(np.r_[0, Ks[:-1] ** 2])
        for k in np.flatnonzero(sreal):
            n = int(Ks[k])
            idx = np.arange(n * n)
            cols, rows = idx // n, idx % n
            ii.append(nb_off + np.maximum(rows, cols) + np.minimum(rows, cols) * n)
            jj.append(jstrt_all[k] + idx)
            vv.append(np.ones(n * n))
            nb_off += n * n

    Kn = {}
    Kn["f"] = 0
    Kn["l"] = newL + 1                      # +1: artificial x0 (pretransfo.m:466-468)
    Kn["q"] = np.r_[newQ, Kq, Kr].astype(np.float64)
    Kn["r"] = np.zeros(0)
    Kn["s"] = Ks[sreal].astype(np.float64)
    Kn["rsdpN"] = int(sreal.sum())
    Kn["N"] = int(Kn["l"] + Kn["q"].sum() + (Kn["s"] ** 2).sum())
    Kn["m"] = m
    Kn["cdim"] = 0

    ii = np.concatenate(ii) + 1 if ii else np.zeros(0, dtype=np.int64)
    jj = np.concatenate(jj) if jj else np.zeros(0, dtype=np.int64)
    vv = np.concatenate(vv) if vv else np.zeros(0)
    QR = sp.csc_matrix((vv, (ii, jj)), shape=(Kn["N"], c.size))
    At2 = sp.csc_matrix(QR @ At)
    At2.sum_duplicates()
    At2.eliminate_zeros()
    At2.sort_indices()
    c2 = np.asarray(QR @ c).ravel()
    finish_K(Kn)
    return At2, b, c2, Kn, QR


def finish_K(K: dict) -> dict:
    """Detailed cone description, pretransfo.m:526-542 (1-based index fields)."""
    q = np.asarray(K["q"], dtype=np.float64).ravel()
    s = np.asarray(K["s"], dtype=np.float64).ravel()
    nr = int(K.get("rsdpN", len(s)))
    Ksr, Ksc = s[:nr], s[nr:]
    K["q"], K["s"], K["rsdpN"] = q, s, nr
    K["blkstart"] = np.cumsum(np.r_[K["l"] + 1, len(q), q - 1, Ksr ** 2, 2 * Ksc ** 2]).astype(np.float64)
    K["rLen"] = float(Ksr.sum())
    K["hLen"] = float(Ksc.sum())
    K["qMaxn"] = float(max([0.0, *q]))
    K["rMaxn"] = float(max([0.0, *Ksr]))
    K["hMaxn"] = float(max([0.0, *Ksc]))
    K["mainblks"] = K["blkstart"][np.cumsum([1, 1, len(q)]) - 1].copy()
    K["qblkstart"] = K["blkstart"][1:2 + len(q)].copy()
    K["sblkstart"] = K["blkstart"][1 + len(q):].copy()
    K["lq"] = float(K["mainblks"][-1] - 1)
    if "N" not in K:
        K["N"] = int(K["l"] + q.sum() + (Ksr ** 2).sum() + 2 * (Ksc ** 2).sum())
    return K


def K_for_mex(K: dict) -> dict:
    """The struct handed to plugins (what conepars reads, sdmauxCone.c:48-134)."""
    keys = ["f", "l", "q", "r", "s", "rsdpN", "rLen", "hLen", "qMaxn", "rMaxn", "hMaxn",
            "blkstart", "mainblks", "qblkstart", "sblkstart", "lq", "N", "m"]
    out = {}
    for k in keys:
        if k in K:
            v = K[k]
            out[k] = np.asarray(v, dtype=np.float64).reshape(1, -1) if np.ndim(v) else float(v)
    return out


def psd_dims(K: dict):
    """(n_k list, 0-based start offsets of each PSD block in x, lenud)."""
    s = np.asarray(K["s"], dtype=np.int64).ravel()
    start = int(K["mainblks"][2]) - 1
    offs = start + np.cumsum(np.r_[0, s[:-1] ** 2]) if len(s) else np.zeros(0, dtype=np.int64)
    return s, np.asarray(offs, dtype=np.int64), int((s ** 2).sum())
