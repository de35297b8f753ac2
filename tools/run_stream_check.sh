set -x
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_fused_pose.py > gpurun_out/bench_fused_pose.json 2> gpurun_out/bench_fused_pose.err; echo rc=$?; tail -2 gpurun_out/bench_fused_pose.json
