#!/usr/bin/env python
"""bench.py -- IPM iterations/s over SeDuMi's normal-equations hot path on B200.

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line.
A "step" is one pass of the per-iteration recipe (sedumi.m:442-466 +
wrapPcg.m:56-59) over the frozen state of the workload:

    invcholfac -> getada1 -> getada2 -> getada3 -> blkchol -> 4 x (fwblkslv, ./d, bwblkslv) -> 12 x psdscale

Workload at every N: BASELINE.json configs[1] = control07 (K.s=[70,35], m=666, dense ADA).  Its
two PSD blocks do not shard, so N>1 runs N independent replicas (one per GPU, no data-path
collective) and reports weak scaling.  `value` = device-resident iterations/s (inputs already in
HBM, CUDA events on the library stream); `e2e` = the same recipe through the reference-facing MEX
plugins with host buffers (host<->device copies inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NSOLVE, NPSD = 4, 12
FRAMES = (np.zeros(0), np.zeros(0))
LOR = None
METRIC = "IPM iterations/sec (ADA'+Cholesky+solve) over the hot-path recipe"


def load_workload(name):
    from sedumi_b200.host import cones, problems, setup
    if name == "control07":
        raw = problems.load_fixture("control07")
    elif name == "arch0":
        raw = problems.load_fixture("arch0")
    elif name == "nb":
        raw = problems.load_fixture("nb")           # BASELINE configs[2]: 793 Lorentz cones, no PSD block
    elif name == "blockdiag64":
        raw = problems.synth_blockdiag_sdp()
    elif name == "maxcut4000":
        raw = problems.synth_maxcut()
    elif name == "maxcut1000":
        raw = problems.synth_maxcut(n=1000, p=0.02)
    elif name == "blockdiag_small":
        raw = problems.synth_blockdiag_sdp(nblk=8, n=60, m=400, nlink=16, density=0.03)
    else:
        raise SystemExit(f"unknown workload {name}")
    At, b, c, K = cones.pretransfo(*raw)[:4]
    perm = np.arange(At.shape[1]) if name.startswith("blockdiag") else None
    S = setup.build_setup(At, b, c, K, perm=perm)
    d = problems.scaling(K, "S1", seed=problems.SEED0 + 2)
    rng = np.random.default_rng(problems.SEED0)
    rhs = rng.standard_normal((S.m, 1))
    psd_x = rng.standard_normal(int((np.asarray(K["s"]) ** 2).sum()))
    global FRAMES, LOR
    FRAMES = problems.synth_frames(K["s"])         # (lab, frms): product-form spectral factor of the PSD iterate
    nq, qd = len(K["q"]), int((np.asarray(K["q"]) - 1).sum()) if len(K["q"]) else 0
    LOR = (rng.standard_normal(nq), rng.standard_normal(qd), rng.standard_normal(S.m), 1e-17 * rng.standard_normal(S.m),
           rng.standard_normal(S.m))                # mu, x(norm-bound part), residual hi/lo/increment for the Lorentz streams
    return S, d, rhs, psd_x


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            t = [x.strip() for x in line.split(",")]
            if len(t) < 9:
                continue
            try:
                sm.append(float(t[1])); mx.append(float(t[2]))
            except ValueError:
                continue
            for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        return {}


# ----------------------------------------------------------------------------- algorithmic work
def kernel_work(S, name):
    """Algorithmic flops / bytes per launch group of the named kernel over ONE iteration."""
    K = S.K
    s = np.asarray(K["s"], dtype=np.int64)
    m = S.m
    nnzL = S.L["L"].nnz
    if name == "gemm_nt_kernel":
        # getada3's W = D(:,R) T (lower triangle): 2 * n(n+1)/2 * r per (constraint, block) pair, plus
        # invcholfac / psdscale products; dominated by getada3 -- count that part exactly
        At = S.At
        start = int(K["mainblks"][2]) - 1
        bs = start + np.r_[0, np.cumsum(s ** 2)]
        fl = 0.0
        for j in range(m):
            rows = At.indices[int(S.Ablkjc[j, 2]):At.indptr[j + 1]]
            for k, n in enumerate(s):
                sel = rows[(rows >= bs[k]) & (rows < bs[k + 1])] - bs[k]
                if sel.size:
                    r = np.unique(np.r_[sel % n, sel // n]).size
                    fl += n * (n + 1) * r
        n3 = float((s ** 3).sum())
        # invcholfac (n^3/3 MACs), psdscale (two triangular products, n^3 flops each), psdinvjmul (two full 2n^3
        # + two lower n^3 products), 2 x psdframeit (lower, n^3), compact-WY accumulation of Q for large blocks
        fl += n3 / 3 * 2 + NPSD * 2 * n3 + 6 * n3 + 2 * n3
        if s.size and s.max() > 2300:
            fl += 3 * 4.0 / 3.0 * n3
        return dict(bound="tensor", work=fl, unit="TFLOP/s")
    if name in ("trail_kernel", "diag_kernel", "trsm_kernel", "factor_small_kernel", "dense_ldl_kernel"):
        cj = np.diff(S.L["L"].indptr) - 1
        return dict(bound="tensor", work=float((cj * (cj + 1)).sum()), unit="TFLOP/s")
    if name == "psdscale_small_kernel":
        return dict(bound="tensor", work=NPSD * 4.0 * float((s ** 3).sum()), unit="TFLOP/s")
    if name == "ada3_dots_kernel":
        # every stored PSD coefficient of A_i is paired with W_j for each j >= i: 20 B per term (index, value, W gather)
        nnz_psd = int((S.At.indptr[1:] - S.Ablkjc[:, 2]).sum())
        by = 20.0 * nnz_psd * (m + 1) / 2 + 8.0 * S.ADA.nnz
        return dict(bound="hbm", work=by, unit="GB/s")
    if name in ("fwsolve_kernel", "bwsolve_kernel", "dense_solve_kernel<fw>", "dense_solve_kernel<bw>"):
        return dict(bound="hbm", work=NSOLVE * (8.0 * nnzL + 24.0 * m), unit="GB/s")
    return dict(bound="hbm", work=8.0 * (S.ADA.nnz + nnzL), unit="GB/s")


def f64_gemm_peak(torch, dev):
    """Measured FP64 GEMM rate (TFLOP/s) on this GPU: torch.matmul f64 4096^3, best of 5."""
    n = 4096
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def run_reference(S, d, rhs, psd_x, steps, warmup):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refpath
    R = refpath.RefHotPath(S)
    for _ in range(warmup):
        R.iteration(d, rhs, psd_x, NSOLVE, NPSD)
    t_mex0 = R.mex.mex_seconds()
    t0 = time.perf_counter()
    t_np = 0.0
    for _ in range(steps):
        t1 = time.perf_counter()
        m0 = R.mex.mex_seconds()
        udsqr, ADA, absd = R.assemble(d)
        if not len(S.K["s"]):                        # getada.m restated in numpy: count it (minus the MEX time inside it)
            t_np += (time.perf_counter() - t1) - (R.mex.mex_seconds() - m0)
        L = R.factor(ADA, absd)
        for _ in range(NSOLVE):
            R.solve(L, rhs)
        if len(S.K["s"]):
            t1 = time.perf_counter()
            for i in range(NPSD):
                ps = R.psdscale(d, psd_x, i & 1)
            t_np += time.perf_counter() - t1
            R.scaling_tail(d, FRAMES[0], FRAMES[1], np.asarray(ps).ravel())
        else:
            R.lorentz_streams(d, LOR[0], LOR[1], LOR[2], LOR[3], LOR[4])
    wall = time.perf_counter() - t0
    t_mex = R.mex.mex_seconds() - t_mex0
    # time inside the reference's mexFunctions + the restated M pieces; harness marshalling excluded
    t = t_mex + t_np
    return dict(seconds=t, wall=wall, steps=steps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="control07")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels one by one instead of replaying a CUDA graph")
    ap.add_argument("--shard", action="store_true",
                    help="strong scaling: shard PSD blocks over the ranks, one NCCL all-reduce of ADA per iteration "
                         "(needs a multi-block workload, e.g. --workload blockdiag64)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{args.workload} (BASELINE.json configs[1]: K.s=[70,35], m=666, dense ADA, 1 supernode)"
              if args.workload == "control07" else args.workload,
              "recipe": f"invcholfac,getada1,getada2,getada3,blkchol,{NSOLVE}x(fwblkslv,./d,bwblkslv),{NPSD}xpsdscale,"
                        "psdinvjmul,2xpsdframeit,urotorder,givensrot (SURVEY 8d)",
              "scaling_state": "S1 mid-run NT scaling (SURVEY 8d), seed 20260926", "parallelism": f"replicas x{args.gpus}",
              "l2": "L2 flushed (256 MiB write) between timed iterations",
              "launch": "one CUDA graph per iteration" if not args.no_graph else "stream launches"}

    if args.impl == "reference":
        if rank != 0:
            return
        S, d, rhs, psd_x = load_workload(args.workload)
        r = run_reference(S, d, rhs, psd_x, args.steps, args.warmup)
        v = r["steps"] / r["seconds"]
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "iterations/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / r["steps"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic scaling on the control07 fixture",
                "config": config,
                "cpu_baseline": {"value": v, "unit": "iterations/s", "cores": 1, "kind": "reference",
                                 "sample": f"{r['steps']} full iterations of the recipe; time inside the reference mexFunctions "
                                           f"(oracle/_ref, gcc -O2) + numpy psdscale; harness marshalling excluded (wall {r['wall']:.2f}s)"},
                "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    from sedumi_b200 import device as sbdev
    S, d, rhs, psd_x = load_workload(args.workload)
    shard_dist = None
    if args.shard and world > 1:
        from sedumi_b200.host import shard as hshard
        owned = hshard.partition_blocks(S.K["s"], world)[rank]
        s_all = np.asarray(S.K["s"], dtype=np.int64)
        xo = np.r_[0, np.cumsum(s_all ** 2)]
        S, d = hshard.shard_compact(S, d, owned, rank)      # owner-computes: this rank's PSD blocks only
        psd_x = np.concatenate([psd_x[xo[k]:xo[k + 1]] for k in owned]) if owned else np.zeros(0)
        global FRAMES
        from sedumi_b200.host import problems as _pb
        FRAMES = _pb.synth_frames(S.K["s"], seed=_pb.SEED0 + 7 + rank)
        shard_dist = dist
        args.no_graph = True                    # the collective is issued by torch.distributed, outside our graph
        config["launch"] = "stream launches"
        config["parallelism"] = (f"PSD blocks sharded over {world} ranks (owner computes: invcholfac, getada3, psdscale, frames, "
                                 "rotations), 1 all-reduce(ADA,absd)/iteration, factor+solves replicated")
    hp = sbdev.HotPath(S, device=local_rank)
    if shard_dist is not None and len(np.asarray(S.L["xsuper"]).ravel()) - 1 > 2:
        info = hp.shard_factor_setup(world, rank)           # elimination-tree subtrees per rank, replicated top
        config["parallelism"] = (f"PSD blocks and etree subtrees sharded over {world} ranks (owner computes); collectives per iteration: "
                                 f"all-reduce(ADA,absd), all-reduce(top fronts, {info['top_len']} doubles), 2 small all-reduces per solve")
    lib = sbdev.lib()
    stream = hp.stream()
    dev = hp.dev
    with torch.cuda.stream(stream):
        hp.set_scaling(d)
        hp.set_rhs(rhs)
        hp.psd_x[:psd_x.size].copy_(torch.from_numpy(psd_x))
        hp.set_frames(*FRAMES)
        if hp.nq:
            for dst, src in ((hp.q_mu, LOR[0]), (hp.q_x, LOR[1]), (hp.r_hi, LOR[2]), (hp.r_lo, LOR[3]), (hp.r_y, LOR[4])):
                dst[:src.size].copy_(torch.from_numpy(src))
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        stream.synchronize()

        def barrier():
            stream.synchronize()
            if world > 1:
                dist.barrier()
            stream.synchronize()

        for _ in range(args.warmup):
            hp.iteration(NSOLVE, NPSD, shard_dist)
        # the iteration is latency-bound at this size: replay it as one CUDA graph
        run_iter = hp.capture(NSOLVE, NPSD) if not args.no_graph else (lambda: hp.iteration(NSOLVE, NPSD, shard_dist))
        for _ in range(2):
            run_iter()
        barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        l0 = lib.sb200_kernel_launches()
        barrier()
        t_wall0 = time.perf_counter()
        for k in range(args.steps):
            flush.fill_(k & 255)                      # evict L2 between timed iterations (untimed)
            ev[k][0].record(stream)
            run_iter()
            ev[k][1].record(stream)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        launches = lib.sb200_kernel_launches() - l0
        if not args.no_graph:
            launches = hp.launches_per_iteration * args.steps      # kernels inside the replayed graphs
        ms = sum(a.elapsed_time(b) for a, b in ev)
        # also a back-to-back (no flush) figure for context
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            run_iter()
        e1.record(stream)
        stream.synchronize()
        ms_warm = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())

        # ---- per-kernel timing pass (CUDA events after every launch on the library stream)
        roof = None
        if rank != 0 and shard_dist is not None:      # the collective needs every rank in the profiling pass too
            for _ in range(args.steps):
                hp.iteration(NSOLVE, NPSD, shard_dist)
            stream.synchronize()
        if rank == 0:
            import ctypes as C
            lib.sb200_prof_begin()
            for _ in range(args.steps):
                hp.iteration(NSOLVE, NPSD, shard_dist)
            buf = C.create_string_buffer(1 << 16)
            lib.sb200_prof_end(buf, C.c_int64(len(buf)))
            prof = {}
            for ln in buf.value.decode().splitlines():
                nm, cnt, tot = ln.split()
                prof[nm] = (int(cnt), float(tot))
            tot_ms = sum(v[1] for v in prof.values())
            top = max(prof.items(), key=lambda kv: kv[1][1])
            nm, (cnt, tms) = top
            w = kernel_work(S, nm)
            per_iter_ms = tms / args.steps
            peaks = measured_peaks()
            if w["bound"] == "hbm":
                achieved = w["work"] / (per_iter_ms * 1e-3) / 1e9
                peak = peaks.get("hbm_gbs", 6650.0)
                src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
            else:
                achieved = w["work"] / (per_iter_ms * 1e-3) / 1e12
                with torch.cuda.stream(torch.cuda.default_stream()):
                    peak = f64_gemm_peak(torch, dev)
                src = ("FP64 (bound=tensor means the FP64 FMA/DMMA pipes): measured in-run, torch.matmul f64 4096^3 best of 5 "
                       "-- MEASURED_PEAKS.json has no FP64 entry")
            traffic = None
            try:        # DRAM bytes per launch of that kernel from the committed ncu --set full capture
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic_r01.json"))).get(args.workload, {}).get(nm)
            except (OSError, ValueError):
                pass
            roof = {"kernel": nm, "bound": w["bound"], "achieved": achieved, "peak": peak, "unit": w["unit"],
                    "frac": achieved / peak, "traffic": traffic, "peak_source": src,
                    "launches_per_step": cnt / args.steps, "ms_per_step_in_kernel": per_iter_ms,
                    "share_of_step": tms / tot_ms,
                    "kernel_ms_per_step": {k: v[1] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])}}

    value = (1 if shard_dist is not None else world) * args.steps / (ms_max * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": "iterations/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "strong" if shard_dist is not None else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic scaling/rhs on the control07 fixture", "config": config,
            "gpu_launches": int(launches), "ms_per_step_no_flush": ms_warm / args.steps, "wall_s": t_wall}
    # ---- e2e: same recipe through the MEX plugins with host buffers; every rank drives its own GPU, max over ranks
    e2e = None
    if not args.no_e2e and shard_dist is None:
        if world > 1:
            dist.barrier()
        e2e = run_e2e(S, d, rhs, psd_x, min(40, max(3, args.steps // 3)), 1)
        if world > 1:
            tt = torch.tensor([e2e["seconds"]], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e["value"] = world * e2e["steps"] / float(tt.item())
            e2e["timed"] += f"; {world} ranks concurrently, slowest rank"
    if rank == 0:
        line["clocks"] = clocks
        line["roofline"] = roof
        if e2e is not None:
            line["e2e"] = e2e
        if not args.no_cpu_baseline and world == 1:
            nb = 60 if args.workload == "control07" else 3
            r = run_reference(S, d, rhs, psd_x, nb, 1)
            line["cpu_baseline"] = {"value": r["steps"] / r["seconds"], "unit": "iterations/s", "cores": 1, "kind": "reference",
                                    "sample": f"{nb} full iterations of the same recipe on the same inputs; reference mexFunctions "
                                              f"(oracle/_ref, gcc -O2, single-threaded like the reference) + numpy psdscale; "
                                              f"harness marshalling excluded (wall {r['wall']:.1f}s)"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_e2e(S, d, rhs, psd_x, steps, world):
    """The recipe through the reference-facing MEX plugins: host numpy in, host numpy out."""
    import scipy.sparse as sp
    from sedumi_b200.host import setup as hsetup
    from sedumi_b200.mx import MexDir
    gpu = MexDir(os.path.join(ROOT, "sedumi_b200", "mex"))
    Km = S.Kmex()
    ADA0 = sp.csc_matrix((np.zeros(S.ADA.nnz), S.ADA.indices, S.ADA.indptr), shape=S.ADA.shape)
    dstruct = {"l": d["l"], "det": d["det"]}
    Lm = hsetup.L_for_mex(S.L)
    pars = {"canceltol": 1e-12, "maxu": 5e5, "abstol": 1e-20}
    nq = len(S.K["q"])
    DAt = {"q": sp.csc_matrix((nq, S.m))}
    xfull = np.r_[np.zeros(1), psd_x]
    lenud = int((np.asarray(S.K["s"]) ** 2).sum())
    sumn = int(np.asarray(S.K["s"]).sum())

    qbs = S.K["qblkstart"].reshape(1, -1)

    def step():
        if nq:                                       # getDAtm.m:40-43 through the ddot plugin
            tr = hsetup.extractA(S.At, S.Ablkjc, 1, 2, int(S.K["mainblks"][0]), int(S.K["mainblks"][1]))
            DAt["q"] = sp.csc_matrix(sp.diags(d["q1"]) @ tr + gpu.ddot(d["q2"], S.At, qbs, S.Ablkjc))
        ud = gpu.invcholfac(d["u"], Km, d["perm"])
        A1 = gpu.getada1(ADA0, S.At, S.Ablkjc[:, 2], S.Aord["lqperm"], dstruct, S.K["qblkstart"].reshape(1, -1))
        A2 = gpu.getada2(A1, DAt, S.Aord, Km)
        A3, absd = gpu.getada3(A2, S.At, S.Ablkjc[:, 2], S.Aord, ud, Km, nlhs=2)
        LL, Ld, sk, ad = gpu.blkchol(Lm, A3, pars, absd, nlhs=4)
        Lf = dict(Lm, L=LL)
        for _ in range(NSOLVE):
            p = gpu.fwblkslv(Lf, rhs)
            y = gpu.bwblkslv(Lf, p / Ld)
        for i in range(NPSD if lenud else 0):
            ps = gpu.psdscale({"u": d["u"], "perm": d["perm"]}, xfull, Km, float(i & 1))
        if not lenud and nq:
            for _ in range(6):
                gpu.qblkmul(LOR[0], LOR[1], qbs)
            for _ in range(3):
                gpu.ddot(d["q2"], LOR[1], qbs)
            gpu.quadadd(LOR[2], LOR[3], LOR[4], nlhs=2)
        if lenud:
            gpu.psdinvjmul(FRAMES[0], FRAMES[1], ps, Km)
            f = gpu.psdframeit(FRAMES[0], FRAMES[1], Km)
            f = gpu.psdframeit(FRAMES[0], FRAMES[1], Km)
            u2, p2, gjc, g = gpu.urotorder(d["u"], Km, 1.1, nlhs=4)
            gpu.givensrot(gjc, g, f, Km)
        return y

    step()
    t0 = gpu.mex_seconds()
    w0 = time.perf_counter()
    for _ in range(steps):
        step()
    t = gpu.mex_seconds() - t0
    wall = time.perf_counter() - w0
    nA, nL, m = S.ADA.nnz, S.L["L"].nnz, S.m
    h2d = 8 * (lenud + (S.K["l"]) + 3 * nA + S.At.nnz * 0 + nA + m + NSOLVE * 2 * (nL + m) + NPSD * 2 * lenud + lenud
               + (2 * lenud + sumn) + 2 * (lenud + sumn) + lenud + 2 * lenud)      # psdinvjmul, 2 psdframeit, urotorder, givensrot
    d2h = 8 * (lenud + 3 * nA + m + nL + 3 * m + NSOLVE * 2 * m + NPSD * lenud + lenud + 2 * lenud + (lenud + 2 * sumn) + lenud)
    return {"value": world * steps / t, "seconds": t, "unit": "iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
            "timed": "time inside the plugins' mexFunction (host numpy buffers in/out, all H2D/D2H inside), "
                     f"Python marshalling of mxArrays excluded; wall incl. marshalling {wall / steps * 1e3:.2f} ms/step",
            "steps": steps}


if __name__ == "__main__":
    main()
