#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bin_major or roi_align" > gpurun_out/t_b.log 2>&1; echo "rc=$?" >> gpurun_out/t_b.log
tail -3 gpurun_out/t_b.log
for d in 0 1 5; do echo "== dbg $d"; TRACE_LAYOUT=rhwc timeout -s KILL 120 python tools/trace_roi_bwd.py $d 2>&1 | head -6; done > gpurun_out/trace_dbg_rhwc.txt
TRACE_LAYOUT=rhwc timeout -s KILL 120 python tools/trace_roi_bwd.py > gpurun_out/trace_bwd_rhwc.txt 2>&1
