#!/bin/bash
# Final GPU session, last part: parity tests of the touched path, re-capture of the top kernel, bench line + reference arm.
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ada_gpu.py tests/test_device_path_gpu.py tests/test_herm_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -2 > gpurun_out/final_tests_d.log; cat gpurun_out/final_tests_d.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ada3_fused -s 2 -c 1 -f -o gpurun_out/top \
    python bench.py --steps 1 --warmup 3 --no-graph --no-e2e --no-cpu-baseline --no-parity --no-secondary > gpurun_out/top.log 2>&1
timeout 900 python bench.py --steps 30 --warmup 5 2> gpurun_out/bench.err | tail -1 > gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>> gpurun_out/bench.err | tail -1 > gpurun_out/bench_ref.json
cat gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'], d['parity']['ok'])"
