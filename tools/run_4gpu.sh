set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --no-cpu-baseline > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_4gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['e2e']['sync_call']['value'], d['clocks'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 tools/bench_c4_stream.py 3 > gpurun_out/bench_c4_stream_4gpu.json 2> gpurun_out/bench_c4_stream_4gpu.err; echo rc=$?
tail -1 gpurun_out/bench_c4_stream_4gpu.json
