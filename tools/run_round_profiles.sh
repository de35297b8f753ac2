set -x
timeout 900 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python tools/profile_run.py 64 2 > gpurun_out/p.log 2>&1 && ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python tools/profile_run.py 64 1 > gpurun_out/ncu.log 2>&1
tail -c 600 gpurun_out/bench_1gpu.json; tail -c 400 gpurun_out/bench_ref.json
