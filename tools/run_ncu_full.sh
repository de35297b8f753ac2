# usage: run_ncu_full.sh <kernel-regex> <count> <out-name>
python tools/profile_run.py 64 2 > gpurun_out/p.log 2>&1 && ncu --set full --import-source on --clock-control none -k "regex:$1" -c $2 -o gpurun_out/$3 -f python tools/profile_run.py 64 1 > gpurun_out/ncu.log 2>&1; tail -2 gpurun_out/ncu.log | cut -c1-200
