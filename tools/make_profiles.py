"""Turn the raw artifacts of one GPU session (gpurun_out/) into the committed summaries under profiles/.

Expected inputs (produced by tools/profile_on_gpu.sh on the B200 box):
  gpurun_out/bench.json          bench.py line (full: e2e + cpu_baseline)
  gpurun_out/bench_ref.json      bench.py --impl reference line
  gpurun_out/launches.csv        ncu --metrics gpu__time_duration.sum launch list of bench.py --no-graph
  gpurun_out/top.ncu-rep         ncu --set full capture of the top kernels (one iteration)
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"


def short(name):
    m = re.match(r"(?:void )?(?:sb::)?(?:\(anonymous namespace\)::)?(\w+)(<[^>]*>)?", name)
    if not m:
        return name
    base, targ = m.group(1), m.group(2) or ""
    if base == "dense_solve_cluster_kernel" or base == "dense_solve_kernel":
        return "dense_solve_kernel<bw>" if ("true" in targ or "1" in targ) else "dense_solve_kernel<fw>"
    return base


def launches():
    path = os.path.join(SRC, "launches.csv")
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[mi] != "gpu__time_duration.sum":
            continue
        k = short(r[ki])
        if k in ("cutlass", "at", "vectorized_elementwise_kernel", "elementwise_kernel"):
            continue        # torch's own kernels: the L2 flush fill and the FP64 matmul that measures the peak
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    return agg


def main():
    os.makedirs(OUT, exist_ok=True)
    bench = json.loads(open(os.path.join(SRC, "bench.json")).read().strip().splitlines()[-1])
    json.dump(bench, open(os.path.join(OUT, f"bench_{TAG}.json"), "w"), indent=1)
    refp = os.path.join(SRC, "bench_ref.json")
    if os.path.exists(refp):
        json.dump(json.loads(open(refp).read().strip().splitlines()[-1]), open(os.path.join(OUT, f"bench_{TAG}_reference_arm.json"), "w"), indent=1)
    agg = launches()
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"launches_{TAG}.csv"), "w") as f:
        f.write("kernel,launches,total_ns,share\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{c},{t:.0f},{t / tot:.4f}\n")
    # ---- shares table: events (bench) vs ncu
    ev = dict(bench["roofline"]["kernel_ms_per_step"])
    ev.pop("nccl_allreduce", None)
    # psdscale and the psdframeit/psdinvjmul congruences are the same kernel under two profiling labels
    fused = ev.pop("psdscale_small_kernel", 0.0) + ev.pop("small_congruence_kernel", 0.0)
    if fused:
        ev["psdscale_small_dmma_kernel"] = fused
    evtot = sum(ev.values())
    lines = ["| kernel | events: ms/step | share | ncu launch list: share |", "|---|---|---|---|"]
    for k, ms in sorted(ev.items(), key=lambda kv: -kv[1]):
        n = agg.get(k, [0, 0.0])[1] / tot if tot else 0.0
        lines.append(f"| `{k}` | {ms:.3f} | {100 * ms / evtot:.1f}% | {100 * n:.1f}% |")
    open(os.path.join(OUT, f"shares_{TAG}.md"), "w").write("\n".join(lines) + "\n")
    # ---- full capture: raw page -> per-kernel text + DRAM traffic
    reps = [os.path.join(SRC, f) for f in ("top.ncu-rep", "top2.ncu-rep") if os.path.exists(os.path.join(SRC, f))]
    if reps:
        keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
                "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
                "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
                "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
                "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"]
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        seen, traffic, txt = {}, {}, []
        for rep in reps:                      # every report carries its own units row
            raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
            rows = list(csv.reader(io.StringIO(raw)))
            hdr, units = rows[0], rows[1]
            ki = hdr.index("Kernel Name")
            idx = {h: i for i, h in enumerate(hdr)}
            for r in rows[2:]:
                k = short(r[ki])
                if k in seen:
                    continue
                seen[k] = 1
                txt.append(f"== {k}   ({r[ki][:100]})   [{os.path.basename(rep)}]")
                for h in keep:
                    if h in idx:
                        txt.append(f"   {h:90s} {r[idx[h]]} {units[idx[h]]}")
                try:
                    rd = float(r[idx['dram__bytes_read.sum']].replace(",", "")) * unit.get(units[idx['dram__bytes_read.sum']], 1.0)
                    wr = float(r[idx['dram__bytes_write.sum']].replace(",", "")) * unit.get(units[idx['dram__bytes_write.sum']], 1.0)
                    traffic[k] = rd + wr
                except (KeyError, ValueError):
                    pass
        open(os.path.join(OUT, f"ncu_{TAG}_top_kernels.txt"), "w").write("\n".join(txt) + "\n")
        wl = bench["config"]["workload"]
        key = "blockdiag64" if "64 PSD blocks" in wl else ("control07" if "control07" in wl else wl)      # bench.py's workload name
        json.dump({key: traffic}, open(os.path.join(OUT, f"traffic_{TAG}.json"), "w"), indent=1)
    print("profiles refreshed:", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
