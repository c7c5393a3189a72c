#!/bin/bash
# Final GPU session, part A (run under gpurun): ncu launch list + ncu --set full captures.  Keep gpurun_out/ below 64 MiB.
set -u
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-parity --no-secondary > gpurun_out/launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ada3_fused -s 2 -c 1 -f -o gpurun_out/top \
    python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-parity --no-secondary > gpurun_out/top.log 2>&1
timeout 400 ncu --set full --clock-control none \
    -k regex:"gemm_nt|wy_rows|factor_small|fwsolve|bwsolve|update_gather|schur_kernel|perm_cols" -s 150 -c 14 -f -o gpurun_out/top2 \
    python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-parity --no-secondary > gpurun_out/top2.log 2>&1
tail -2 gpurun_out/top.log; tail -2 gpurun_out/top2.log; ls -la gpurun_out/
