#!/bin/bash
# round 2: band-ordered cluster passes -- parity first (every quad-level test), then A/B bench against the tile passes, then ncu
python -m pytest tests/test_gpu_detector.py tests/test_gpu_pipeline.py -m gpu -x -q > gpurun_out/r2_bands_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_bands_pytest.log
for v in bands tiles; do
  CB_CLUSTERS=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c4 --no-sqpnp > gpurun_out/r2_cl_$v.json 2> gpurun_out/r2_cl_$v.err || tail -3 gpurun_out/r2_cl_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r2_cl_$v.json'))
print('$v', 'c1 value', round(d['value']), {k:round(x,3) for k,x in d['stage_ms_per_step'].items()})
print('$v', 'c2 value', round(d['also_c2']['value']), {k:round(x,3) for k,x in d['also_c2']['stage_ms_per_step'].items()})
PY
done
python tools/profile_run.py 256 2 c1 > gpurun_out/p.log 2>&1 && ncu --set full --import-source on --clock-control none -k "regex:threshold_tm" -c 1 -o gpurun_out/r2_thr_tm -f python tools/profile_run.py 256 1 c1 > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log | cut -c1-200
python tools/ncu_summary.py gpurun_out/r2_thr_tm.ncu-rep > gpurun_out/r2_thr_tm_summary.txt 2>&1; head -40 gpurun_out/r2_thr_tm_summary.txt
