#!/bin/bash
# 8-GPU evidence: the H2D ceiling of the box with 8 ranks copying at once, then bench.py at 8 GPUs (weak scaling + the c4 strong-scaling stream)
N=${1:-8}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/h2d_ceiling.py > gpurun_out/r02_h2d_ceiling_${N}gpu.json 2> gpurun_out/r02_h2d_ceiling_${N}gpu.err; echo rc=$?
tail -1 gpurun_out/r02_h2d_ceiling_${N}gpu.json | cut -c1-900
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${N}gpu.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['sync_call']['value'], d['roofline']['frac'], d['clocks']); c=d['c4_stream']; print(c['multi_process']['value'], c['single_process_pool'].get('value'), c['single_process_pool'].get('error'))"
