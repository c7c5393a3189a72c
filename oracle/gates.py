"""Parity gates of SURVEY.md section 8d -- TEST INFRASTRUCTURE (used by tests/ and by bench.py's untimed
in-run parity leg; never by the product path).

`run_gates` replays ONE iteration of the recipe through the reference (oracle/refpath.py: the unmodified
reference MEX targets in oracle/_ref + the numpy restatement of the M-only psdscale) on the same inputs a
`sedumi_b200.device.HotPath` has just processed, and compares what both produced:

    ADA, absd            <= 1e-10 relative (Frobenius)          getada1-3
    L (factor), d        <= 1e-10 (Frobenius / infinity norm)   blkchol; identical skip and add index sets
    y (search direction) <= 1e-8  relative                      fwblkslv, ./d, bwblkslv
    udsqr, psdscale, frames (psdframeit), urotorder (bit-exact u/perm/gjc), givensrot  <= 1e-10

It also reports how long the reference spent inside its mexFunctions (+ numpy psdscale) for that iteration, so
that the same CPU work serves as bench.py's `cpu_baseline` sample.
"""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

import refpath

TOL = {"d_dense": 1e-10, "ADA": 1e-10, "absd": 1e-10, "L": 1e-10, "d": 1e-10, "y": 1e-8, "udsqr": 1e-10, "psdscale": 1e-10,
       "frame": 1e-10, "givensrot": 1e-10}


def _rel(a, b, ord=None):
    a = np.asarray(a, dtype=float).ravel()
    b = np.asarray(b, dtype=float).ravel()
    nb = np.linalg.norm(b, ord)
    return float(np.linalg.norm(a - b, ord) / (nb if nb > 0 else 1.0))


def device_outputs(hp, S, nnzL):
    """What the device holds after hp.iteration(...): host copies."""
    from sedumi_b200 import device
    L = device.lib()
    hp.sync()
    t = hp.torch
    Lcsc = t.empty(max(nnzL, 1), dtype=t.float64, device=hp.dev)
    device.check(L.sb200_chol_rect_to_csc_dev(hp.chol, C.c_void_p(hp.Lrect.data_ptr()), C.c_void_p(hp.flag.data_ptr()),
                                              C.c_void_p(Lcsc.data_ptr())), "rect_to_csc")
    hp.sync()
    n2 = hp.lenud
    out = dict(ADA=hp.ADA.cpu().numpy()[:S.ADA.nnz], absd=hp.absd.cpu().numpy()[:S.m], L=Lcsc.cpu().numpy()[:nnzL],
               d=hp.dvec.cpu().numpy()[:S.m], flag=hp.flag.cpu().numpy()[:S.m], y=hp.y.cpu().numpy().T.copy())
    if n2:
        out.update(udsqr=hp.udsqr.cpu().numpy()[:n2], psd=hp.psd_y.cpu().numpy()[:n2], frame=hp.psd_f.cpu().numpy()[:n2],
                   u=hp.u_new.cpu().numpy()[:n2], perm=hp.perm_new.cpu().numpy()[:hp.sumn], gjc=hp.gjc.cpu().numpy()[:hp.sumn],
                   q=hp.psd_z.cpu().numpy()[:n2])
    return out


def run_gates(hp, S, d, rhs, psd_x, frames, nsolve, npsd, *, S_local=None, d_local=None, colmask=None, tail=True):
    """Compare the state `hp` holds after hp.iteration(nsolve, npsd[, sharded]) with the reference.
    S/d: the FULL problem and scaling (ADA, factor, solve); S_local/d_local: this rank's cone when the PSD part is
    sharded (owner computes), else None; colmask: columns of the factor this rank answers for (sharded) or None.
    Returns {"ok": bool, "err": {...}, "tol": {...}, "skip_equal", "add_equal", "ref_seconds", ...}."""
    Sl = S_local if S_local is not None else S
    dl = d_local if d_local is not None else d
    nnzL = int(S.L["L"].nnz)
    dev = device_outputs(hp, S, nnzL)
    R = refpath.RefHotPath(S)
    t_np = 0.0
    m0 = R.mex.mex_seconds()
    w0 = time.perf_counter()
    udsqr, ADA, absd = R.assemble(d)
    from sedumi_b200.host import setup as hsetup
    LL, Ld, skip, add = R.mex.blkchol(hsetup.L_for_mex(dict(S.L)), ADA, R.pars, absd, nlhs=4)
    Lref = R.factor(ADA, absd)
    y = None
    dense = getattr(S, "dense", None)
    Lden = dden = None
    if dense is not None and len(dense.cols):            # deninfac.m:57-79 on the reference's own MEX files
        DC = refpath.DenseColumnRef(S, Lref)
        LAD, Ld0, sym, smult = DC.inputs(d, np.asarray(Ld, dtype=float).ravel().copy())
        Lden, dden = R.mex.dpr1fact(LAD, Ld0.reshape(-1, 1), sym, smult, 5e2, nlhs=2)
        Lden.update(dz=sym["dz"], first=sym["first"], perm=sym["perm"])
        Lm = hsetup.L_for_mex({k: Lref[k] for k in ("perm", "L", "xsuper", "tmpsiz")})
        for _ in range(nsolve):
            y = R.mex.bwblkslv(Lm, R.mex.bwdpr1(Lden, R.mex.fwdpr1(Lden, R.mex.fwblkslv(Lm, rhs)) / np.asarray(dden).reshape(-1, 1)))
    else:
        for _ in range(nsolve):
            y = R.solve(Lref, rhs)
    err, info = {}, {}
    err["ADA"] = _rel(dev["ADA"], ADA.data)
    err["absd"] = _rel(dev["absd"], absd)
    Ld = np.asarray(Ld, dtype=float).ravel()
    if colmask is None:
        err["L"] = _rel(dev["L"], LL.data)
        err["d"] = _rel(dev["d"], Ld, np.inf)
        sk_dev, ad_dev = np.flatnonzero(dev["flag"] == 1), np.flatnonzero(dev["flag"] == 2)
        info["skip_equal"] = bool(np.array_equal(sk_dev, skip.indices))
        info["add_equal"] = bool(np.array_equal(ad_dev, add.indices))
    else:                                   # sharded factor: the columns this rank answers for (its subtrees + the top)
        cols = np.flatnonzero(colmask)
        jc = LL.indptr
        sel = np.concatenate([np.arange(jc[c], jc[c + 1]) for c in cols]) if cols.size else np.zeros(0, dtype=np.int64)
        err["L"] = _rel(dev["L"][sel], LL.data[sel])
        err["d"] = _rel(dev["d"][cols], Ld[cols], np.inf)
        info["skip_equal"] = bool(np.array_equal(np.flatnonzero(dev["flag"][cols] == 1), np.flatnonzero(np.isin(cols, skip.indices))))
        info["add_equal"] = bool(np.array_equal(np.flatnonzero(dev["flag"][cols] == 2), np.flatnonzero(np.isin(cols, add.indices))))
        info["factor_scope"] = f"{cols.size} of {S.m} columns (this rank's subtrees + replicated top)"
    err["y"] = _rel(dev["y"], y)
    if dden is not None:
        err["d_dense"] = _rel(hp.dvec_den.cpu().numpy()[:S.m], np.asarray(dden).ravel(), np.inf)
    info["nskip"], info["nadd"] = int(skip.nnz), int(add.nnz)
    if hp.lenud:
        Rl = R if Sl is S else refpath.RefHotPath(Sl)
        ud_l = udsqr if Sl is S else Rl.mex.invcholfac(dl["u"], Rl.Km, dl["perm"])
        err["udsqr"] = _rel(dev["udsqr"], ud_l)
        ps = None
        t1 = time.perf_counter()
        for i in range(npsd):
            ps = Rl.psdscale(dl, psd_x, i & 1)
        t_np += time.perf_counter() - t1
        if ps is not None:
            err["psdscale"] = _rel(dev["psd"], ps)
        if tail and ps is not None:
            rt = Rl.scaling_tail(dl, frames[0], frames[1], np.asarray(ps).ravel())
            err["frame"] = _rel(dev["frame"], rt["frame"])
            err["givensrot"] = _rel(dev["q"], rt["q"])
            info["urotorder_bit_exact"] = bool(np.array_equal(dev["u"], rt["u"].ravel()) and
                                               np.array_equal(dev["perm"] + 1, rt["perm"].ravel().astype(np.int64)) and
                                               np.array_equal(dev["gjc"], rt["gjc"].ravel().astype(np.int64)))
    ref_seconds = (R.mex.mex_seconds() - m0) + t_np
    if Sl is not S and hp.lenud:
        ref_seconds += 0.0          # the local reference's MEX time is part of the same MexDir clock (shared directory)
    ok = all(err[k] <= TOL[k] for k in err) and info.get("skip_equal", True) and info.get("add_equal", True) \
        and info.get("urotorder_bit_exact", True)
    return dict(ok=bool(ok), err=err, tol={k: TOL[k] for k in err}, ref_seconds=ref_seconds,
                ref_wall=time.perf_counter() - w0, **info)
