#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bin_major or bench_shape or bench_size" > gpurun_out/t_e.log 2>&1; echo "rc=$?" >> gpurun_out/t_e.log
tail -3 gpurun_out/t_e.log
timeout -s KILL 200 python tools/step_once.py 2 > gpurun_out/step_once.log 2>&1 && \
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_raw.csv python tools/step_once.py 2 > gpurun_out/ncu_step.log 2>&1
echo "launch list rc=$?"
timeout -s KILL 100 python tools/prof_roi2.py > gpurun_out/prof_roi2.log 2>&1 && \
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:roi_align_.*_tc --launch-skip 4 --launch-count 2 -f -o gpurun_out/r02_roi_rhwc python tools/prof_roi2.py > gpurun_out/ncu_roi.log 2>&1
echo "ncu full rc=$?"
ncu -i gpurun_out/r02_roi_rhwc.ncu-rep --page raw --csv > gpurun_out/r02_roi_rhwc_raw.csv 2>/dev/null; echo "raw rc=$?"
ls -la gpurun_out | tail -12
