set -x
CB_THR_VERBOSE=1 timeout 300 python tools/profile_run.py 256 2 c1 > gpurun_out/p.log 2> gpurun_out/r02_thr_plan.txt || exit 1
grep "threshold plan" gpurun_out/r02_thr_plan.txt
export $(grep -o "CB_THR_CFG=[0-9]*" gpurun_out/r02_thr_plan.txt | tail -1) $(grep -o "CB_THR_YSEGS=[1-9][0-9]*" gpurun_out/r02_thr_plan.txt | tail -1)      # (YSEGS=0: the wave model decides, nothing to pin)
echo "pinned: CB_THR_CFG=$CB_THR_CFG CB_THR_YSEGS=$CB_THR_YSEGS"
timeout 600 ncu --set full --import-source on --clock-control none -k "regex:threshold_tm" -c 1 -o gpurun_out/r02_thr -f python tools/profile_run.py 256 2 c1 > gpurun_out/ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_thr.ncu-rep threshold --traffic c1 256 "ncu --set full of the pipeline's threshold launch on 256 x 1280x720 (profiles/r02_ncu_threshold.txt): (dram read + write) / 256 frames" > gpurun_out/r02_ncu_threshold.txt 2>&1; cat gpurun_out/r02_ncu_threshold.txt
cp profiles/threshold_traffic.json gpurun_out/r02_threshold_traffic.json
for i in 1 2 3; do CB_THR_VERBOSE=1 timeout 100 python tools/profile_run.py 256 1 c1 2>&1 | grep "threshold plan"; done
