"""Multi-GPU sharding of the hot path.

The path shards naturally over PSD blocks: every PSD kernel is block-diagonal and ADA is a SUM over
blocks of per-block contributions (getada3.c:333-351 accumulates into one array).  So:

  * PSD block k is owned by one rank (greedy bin-packing on n_k^3);
  * each rank keeps only the rows of At that belong to its blocks (rank 0 also keeps the LP/Lorentz
    rows), runs getada1/2/3 on that shard against the FULL ADA pattern, and
  * ONE all-reduce (sum) of the ADA values and of absd at the Schur-assembly boundary gives every
    rank the complete matrix; the factor and the solves are then replicated.

Both ADA and absd are linear in the per-block terms, and the symmetrisation X+X'-diag is linear too,
so the reduction commutes with everything getada3 does.
"""
from __future__ import annotations

import copy

import numpy as np
import scipy.sparse as sp

from . import setup as hsetup


def partition_blocks(s, world: int) -> list[list[int]]:
    """Greedy longest-processing-time assignment of PSD blocks to ranks, cost n^3."""
    s = np.asarray(s, dtype=np.int64).ravel()
    order = np.argsort(-(s ** 3), kind="stable")
    load = np.zeros(world)
    owned = [[] for _ in range(world)]
    for k in order:
        r = int(np.argmin(load))
        owned[r].append(int(k))
        load[r] += float(s[k]) ** 3
    return [sorted(o) for o in owned]


def shard_setup(S, owned: list[int], rank: int):
    """Copy of the setup whose At holds only this rank's share of the rows."""
    K = S.K
    s = np.asarray(K["s"], dtype=np.int64)
    start = int(K["mainblks"][2]) - 1
    bs = start + np.r_[0, np.cumsum(s ** 2)]
    keep = np.zeros(S.At.shape[0], dtype=bool)
    if rank == 0:
        keep[:start] = True
    for k in owned:
        keep[bs[k]:bs[k + 1]] = True
    At = sp.csc_matrix(sp.diags(keep.astype(np.float64)) @ S.At)
    At.eliminate_zeros()
    At.sort_indices()
    T = copy.copy(S)
    T.At = At
    T.Ablkjc = hsetup.partitA(At, K["mainblks"])
    sperm, dz = hsetup.incorder(At, T.Ablkjc[:, 2], int(K["mainblks"][2]))
    T.Aord = dict(S.Aord, sperm=sperm.reshape(-1, 1), dz=dz)
    return T


def shard_compact(S, d, owned: list[int], rank: int):
    """Owner-computes shard: the cone of this rank keeps only its own PSD blocks -- K.s, the PSD rows of
    At, d.u / d.perm -- so every per-block kernel (invcholfac, getada3's products, psdscale, psdframeit, ...) does
    1/world of the work; rank 0 also keeps the LP/Lorentz rows.  The ADA pattern, the symbolic factor and the
    constraint numbering stay global: the partial ADA values of all ranks add up to the full matrix.
    Returns (setup, scaling) of the shard."""
    K = S.K
    s = np.asarray(K["s"], dtype=np.int64)
    nr = int(K.get("rsdpN", len(s)))
    start = int(K["mainblks"][2]) - 1
    span = np.where(np.arange(len(s)) < nr, 1, 2) * s ** 2
    bs = start + np.r_[0, np.cumsum(span)]
    rows = [np.arange(start)] + [np.arange(bs[k], bs[k + 1]) for k in owned]
    rows = np.concatenate(rows)
    At = sp.csr_matrix(S.At)[rows, :].tocsc()
    if rank != 0:                                   # the LP / Lorentz part is assembled by rank 0 only
        scale = np.ones(rows.size)
        scale[:start] = 0.0
        At = sp.csc_matrix(sp.diags(scale) @ At)
    At.eliminate_zeros()
    At.sort_indices()
    Kr = {"l": K["l"], "q": np.asarray(K["q"]).copy(), "s": s[owned].astype(np.float64),
          "rsdpN": int(sum(1 for k in owned if k < nr))}
    from . import cones
    Kr = cones.finish_K(Kr)
    for key in ("m", "cdim"):
        if key in K:
            Kr[key] = K[key]
    T = copy.copy(S)
    T.At, T.K = At, Kr
    T.Ablkjc = hsetup.partitA(At, Kr["mainblks"])
    sperm, dz = hsetup.incorder(At, T.Ablkjc[:, 2], int(Kr["mainblks"][2]))
    T.Aord = dict(S.Aord, sperm=sperm.reshape(-1, 1), dz=dz)
    # scaling of the owned blocks
    u = np.asarray(d["u"], dtype=np.float64).ravel()
    uo = np.r_[0, np.cumsum(span)]
    po = np.r_[0, np.cumsum(s)]
    dr = dict(d)
    dr["u"] = np.concatenate([u[uo[k]:uo[k + 1]] for k in owned]) if owned else np.zeros(0)
    p = np.asarray(d.get("perm", np.zeros(0))).ravel()
    dr["perm"] = (np.concatenate([p[po[k]:po[k + 1]] for k in owned]).reshape(-1, 1) if (p.size and owned) else np.zeros((0, 0)))
    return T, dr
