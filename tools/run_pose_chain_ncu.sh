# ncu launch list of the pose chain of ONE frame (plain launches: CB_GRAPH=0), after the same command exited 0 without ncu
python tools/profile_pose.py > gpurun_out/pp0.log 2>&1 || exit 1
CB_GRAPH=0 timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k "regex:sq_|assemble" --csv --log-file gpurun_out/pose_launches.csv python tools/profile_pose.py > gpurun_out/pp.log 2>&1
tail -2 gpurun_out/pp.log
