#!/bin/bash
# A/B of the threshold kernels on the GPU box: parity tests that exercise every threshold path, then the bench per variant
python -m pytest tests/test_gpu_detector.py -m gpu -x -q -k "stages_bit_exact or strided or c3_full or edge_cases or golden_fixture or other_decimation" > gpurun_out/thr_pytest.log 2>&1; tail -4 gpurun_out/thr_pytest.log
for v in tmap tma; do
  CB_THRESHOLD=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c4 --no-sqpnp > gpurun_out/thr_$v.json 2> gpurun_out/thr_$v.err || tail -3 gpurun_out/thr_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/thr_$v.json'))
print('$v', 'c1 thr ms', d['roofline']['ms_per_launch'], 'frac', round(d['roofline']['frac'],3), 'value', round(d['value']), '| c2 thr GB/s', round(d['also_c2']['threshold_gbs']), 'c2 value', round(d['also_c2']['value']))
PY
done
for t in 4 6; do for y in 2 3 5 9; do
  CB_THR_T=$t CB_THR_YSEGS=$y python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-c4 --no-sqpnp --no-c2 --latency-iters 1 > gpurun_out/thr_t${t}_y$y.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/thr_t${t}_y$y.json'))
print('T=$t ysegs=$y', 'c1 thr ms', round(d['roofline']['ms_per_launch'],4), 'frac', round(d['roofline']['frac'],3))
PY
done; done
