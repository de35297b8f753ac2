set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 python tools/bench_configs.py 5 2> /dev/null | tail -3 > gpurun_out/bench_configs.jsonl
timeout 600 python bench_sqpnp.py 1000000 2> /dev/null | tail -1 > gpurun_out/bench_sqpnp.json
python tools/profile_run.py 64 2 > gpurun_out/p.log 2>&1 && ncu --set full --import-source on --clock-control none -k "regex:." -c 30 -o gpurun_out/all -f python tools/profile_run.py 64 1 > gpurun_out/ncu.log 2>&1
tail -c 300 gpurun_out/bench_sqpnp.json
