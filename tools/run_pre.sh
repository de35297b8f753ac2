set -x
timeout 600 python -m pytest tests/test_gpu_detector.py tests/test_gpu_cat.py -m gpu -q -x -k "rgb or cat or gray" 2>&1 | tail -5
timeout 300 python tools/bench_preprocess.py > gpurun_out/bench_preprocess.json 2> gpurun_out/bench_preprocess.err; echo rc=$?; cat gpurun_out/bench_preprocess.json; tail -3 gpurun_out/bench_preprocess.err
