# ncu launch list of bench.py itself (the device-resident arm: 3 warm-up + 2 timed steps = the first 120 kernel launches),
# taken only after the same command has exited 0 without ncu.  Numbers printed under ncu are never bench values.
set -x
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c1 --latency-iters 0 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-c1 --latency-iters 0 > gpurun_out/ncu_bench.log 2>&1
python tools/launch_table.py gpurun_out/launches_bench.csv 24 | tee gpurun_out/launches_bench_last_step.txt | tail -30
