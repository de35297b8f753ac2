#!/bin/bash
# last evidence run of round 2 after the pose-chain work (detector kernels unchanged since tools/run_r2_profiles.sh ran): GPU tests,
# smoke, bench (both arms).  Everything lands in gpurun_out/r02_*.
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; echo rc=$?
timeout 600 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err; echo rc=$?
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json'))
print({k:d[k] for k in ('value','ms_per_step','p50_frame_latency_ms','p50_detect_pose_latency_ms')}, d['e2e']['value'], d['roofline']['frac'], d['also_c2']['value'], d['sqpnp_1M'])
"
