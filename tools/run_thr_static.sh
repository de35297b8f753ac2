#!/bin/bash
# in-pipeline sweep of pinned threshold shapes (no shape timing): roofline of c1 and threshold GB/s of c2 per (CB_THR_CFG, CB_THR_YSEGS)
for c in 1 0 2; do for y in 0 5 6 9 12 18; do
  if [ $y = 0 ]; then unset CB_THR_YSEGS; else export CB_THR_YSEGS=$y; fi
  CB_THR_CFG=$c timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-c4 --no-sqpnp --latency-iters 1 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('cfg $c ysegs $y: c1 frac', round(d['roofline']['frac'],3), 'us', round(d['roofline']['ms_per_launch']*1e3,1), '| c2 GB/s', round(d['also_c2']['threshold_gbs']), 'frac', round(d['also_c2']['threshold_gbs']/6533.2,3))"
done; done
