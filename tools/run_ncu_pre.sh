set -x
python tools/profile_preprocess.py 2 > gpurun_out/pp.log 2>&1 || { tail -5 gpurun_out/pp.log; exit 1; }
timeout 200 ncu --set full --import-source on --clock-control none -k "regex:to_gray" -c 2 -o gpurun_out/preprocess -f python tools/profile_preprocess.py 1 > gpurun_out/ncu_pre.log 2>&1; tail -2 gpurun_out/ncu_pre.log | cut -c1-200
timeout 300 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
