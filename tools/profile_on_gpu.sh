#!/bin/bash
# Run on the B200 box (gpurun): bench + reference arm + ncu launch list + one ncu --set full capture.
# Results land in gpurun_out/; tools/make_profiles.py turns them into profiles/.
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --warmup 5 2> gpurun_out/bench.err | tail -1 > gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>> gpurun_out/bench.err | tail -1 > gpurun_out/bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-parity --no-secondary > gpurun_out/launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ada3_fused -s 2 -c 1 -f -o gpurun_out/top \
    python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-parity --no-secondary > gpurun_out/top.log 2>&1
timeout 400 ncu --set full --clock-control none \
    -k regex:"gemm_nt|wy_rows|factor_small|fwsolve|bwsolve|update_gather|schur_kernel|tri_transpose|perm_cols|ada3_reduce|wy_t_kernel|sym_ops" -s 120 -c 48 -f -o gpurun_out/top2 \
    python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-parity --no-secondary > gpurun_out/top2.log 2>&1
tail -2 gpurun_out/top.log
cat gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['kernel'], d['roofline']['frac'])"
cat gpurun_out/bench_ref.json | cut -c1-200
