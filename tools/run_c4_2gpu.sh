set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/bench_c4_stream.py 3 > gpurun_out/bench_c4_stream_2gpu.json 2> gpurun_out/bench_c4_stream_2gpu.err; echo rc=$?
tail -1 gpurun_out/bench_c4_stream_2gpu.json; tail -3 gpurun_out/bench_c4_stream_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo rc=$?
python -c "
import json; d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['e2e']['sync_call']['value'])"
