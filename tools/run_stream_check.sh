set -x
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_stream2.json 2> gpurun_out/bench_stream2.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_stream2.json')); print(d['value'], d['e2e'], d['also_1280x720'], d['clocks'])"
timeout 300 python tools/bench_c4_stream.py 3 > gpurun_out/bench_c4_stream_1gpu.json 2> gpurun_out/bench_c4_stream_1gpu.err; echo rc=$?; cat gpurun_out/bench_c4_stream_1gpu.json; tail -3 gpurun_out/bench_c4_stream_1gpu.err
