for v in 0 2049 1025; do
  CB_LFPS_BIG=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c4 --no-sqpnp --latency-iters 1 > gpurun_out/big_$v.json 2>/dev/null
  python - <<PY
import json
d=json.load(open('gpurun_out/big_$v.json'))
print('CB_LFPS_BIG=$v c1 value', round(d['value']), 'quad', round(d['stage_ms_per_step']['quad_ms'],3), '| c2 value', round(d['also_c2']['value']), 'quad', round(d['also_c2']['stage_ms_per_step']['quad_ms'],3))
PY
done
