set -x
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; echo rc=$?
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_1gpu.json')); print(d['value'], d['e2e']['value'], d['e2e']['sync_call']['value'], d['roofline']['frac'], d['p50_frame_latency_ms'], d['cpu_baseline']['value'], d['clocks'])
r=json.load(open('gpurun_out/bench_ref.json')); print(r['value'], r['impl'])"
