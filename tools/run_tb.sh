timeout 600 python -m pytest tests/test_gpu_detector.py -x -q 2>&1 | tail -3 | tr '\n' ' '; echo
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/b.json; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(round(d['value']), round(d['e2e']['value'])); print({k:round(v,3) for k,v in d['stage_ms_per_step'].items()})"
