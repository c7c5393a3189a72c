"""Per-plugin time of the MEX-boundary (e2e) recipe: where do the milliseconds go?"""
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from sedumi_b200.host import setup as hsetup  # noqa: E402
from sedumi_b200.mx import MexDir  # noqa: E402

S, d, rhs, psd_x = bench.load_workload(sys.argv[1] if len(sys.argv) > 1 else "control07")
gpu = MexDir(os.path.join(ROOT, "sedumi_b200", "mex"))
Km = S.Kmex()
ADA0 = sp.csc_matrix((np.zeros(S.ADA.nnz), S.ADA.indices, S.ADA.indptr), shape=S.ADA.shape)
Lm = hsetup.L_for_mex(S.L)
pars = {"canceltol": 1e-12, "maxu": 5e5, "abstol": 1e-20}
DAt = {"q": sp.csc_matrix((len(S.K["q"]), S.m))}
xfull = np.r_[np.zeros(1), psd_x]


def step():
    ud = gpu.invcholfac(d["u"], Km, d["perm"])
    A1 = gpu.getada1(ADA0, S.At, S.Ablkjc[:, 2], S.Aord["lqperm"], {"l": d["l"], "det": d["det"]}, S.K["qblkstart"].reshape(1, -1))
    A2 = gpu.getada2(A1, DAt, S.Aord, Km)
    A3, absd = gpu.getada3(A2, S.At, S.Ablkjc[:, 2], S.Aord, ud, Km, nlhs=2)
    LL, Ld, sk, ad = gpu.blkchol(Lm, A3, pars, absd, nlhs=4)
    Lf = dict(Lm, L=LL)
    for _ in range(4):
        p = gpu.fwblkslv(Lf, rhs)
        gpu.bwblkslv(Lf, p / Ld)
    for i in range(12):
        ps = gpu.psdscale({"u": d["u"], "perm": d["perm"]}, xfull, Km, float(i & 1))
    F = bench.FRAMES
    gpu.psdinvjmul(F[0], F[1], ps, Km)
    f = gpu.psdframeit(F[0], F[1], Km)
    f = gpu.psdframeit(F[0], F[1], Km)
    u2, p2, gjc, g = gpu.urotorder(d["u"], Km, 1.1, nlhs=4)
    gpu.givensrot(gjc, g, f, Km)


step(); step()
for p in gpu._cache.values():
    p.seconds = 0.0; p.calls = 0
n = 5
t0 = time.perf_counter()
for _ in range(n):
    step()
wall = time.perf_counter() - t0
print(f"wall {wall / n * 1e3:.2f} ms/step, inside mexFunction {gpu.mex_seconds() / n * 1e3:.2f} ms/step")
for name, p in sorted(gpu._cache.items(), key=lambda kv: -kv[1].seconds):
    print(f"  {name:12s} {p.calls // n:3d} calls/step  {p.seconds / n * 1e3:8.3f} ms/step  {p.seconds / max(p.calls, 1) * 1e3:8.3f} ms/call")
