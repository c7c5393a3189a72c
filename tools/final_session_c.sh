set -u
timeout 900 python bench.py --steps 30 --warmup 5 2> gpurun_out/bench.err | tail -1 > gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 2>> gpurun_out/bench.err | tail -1 > gpurun_out/bench_ref.json
cat gpurun_out/bench.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'])"
