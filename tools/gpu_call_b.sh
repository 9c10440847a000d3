#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bin_major or roi_align" > gpurun_out/t_b.log 2>&1; echo "rc=$?" >> gpurun_out/t_b.log
tail -4 gpurun_out/t_b.log
for lay in rhwc; do TRACE_LAYOUT=$lay timeout -s KILL 120 python tools/trace_roi_bwd.py > gpurun_out/trace_bwd_$lay.txt 2>&1; echo "trace $lay rc=$?"; done
timeout -s KILL 200 python bench.py --no-cpu-baseline --no-f32-line --roi-layout rhwc > gpurun_out/b_rhwc.json 2> gpurun_out/b_rhwc.err; echo "rc=$?"
python - <<'P'
import json
for n in ("rhwc",):
    try:
        d=json.loads(open(f"gpurun_out/b_{n}.json").read().strip().splitlines()[-1])
        k=d["kernels"]
        print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], {x:(k[x]["ms"],k[x]["frac"]) for x in k if x.startswith("roi_align")})
    except Exception as e:
        print(n, "failed", e)
P
