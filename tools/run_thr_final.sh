#!/bin/bash
# round 2: threshold kernel with TMA bulk stores / row-major stores + per-geometry shape timing: parity, micro-benchmark table, bench line
timeout 600 python -m pytest tests/test_gpu_detector.py -m gpu -x -q > gpurun_out/thr_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/thr_final_pytest.log
THR_SHORT=1 timeout 60 tools/cuda/thr_bench 1280 720 256 > gpurun_out/thr_bench_1280x720.txt 2>&1; grep -v plain gpurun_out/thr_bench_1280x720.txt | head -40
THR_SHORT=1 timeout 60 tools/cuda/thr_bench 1456 1088 256 > gpurun_out/thr_bench_1456x1088.txt 2>&1
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c4 --no-sqpnp > gpurun_out/thr_final_bench.json 2> gpurun_out/thr_final_bench.err || tail -5 gpurun_out/thr_final_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/thr_final_bench.json'))
print('c1 thr ms', round(d['roofline']['ms_per_launch'],4), 'frac', round(d['roofline']['frac'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), '| c2 thr GB/s', round(d['also_c2']['threshold_gbs']), 'c2 value', round(d['also_c2']['value']))
print(d['stage_ms_per_step'])
PY
