#!/bin/bash
# Run on the B200 box (gpurun): bench + reference arm + ncu launch list + one ncu --set full capture.
# Results land in gpurun_out/; tools/make_profiles.py turns them into profiles/.
set -u
mkdir -p gpurun_out
timeout 300 python bench.py --steps 30 --warmup 5 2> gpurun_out/bench.err | tail -1 > gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>> gpurun_out/bench.err | tail -1 > gpurun_out/bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:"dense_ldl|dense_solve|ada3_dots|psdscale_small|gemm_nt|householder_q|urotorder_kernel|build_tt|makesym" -s 120 -c 40 -f -o gpurun_out/top \
    python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/top.log 2>&1
tail -2 gpurun_out/top.log
cat gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['kernel'], d['roofline']['frac'])"
cat gpurun_out/bench_ref.json | cut -c1-200
