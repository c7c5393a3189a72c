#!/bin/bash
# Final GPU session: the whole GPU suite and the bench line on the final code.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_tests_e.log; cat gpurun_out/final_tests_e.log
timeout 600 python bench.py --steps 30 --warmup 5 2> gpurun_out/bench.err | tail -1 > gpurun_out/bench.json
cat gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], d['cpu_baseline']['value'], d['roofline']['frac'], d['parity']['ok']); print([ (s['workload'][:12], round(s['value'],1), s.get('e2e',{}).get('value')) for s in d['secondary']])"
