#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for lay in rchw rhwc; do
timeout -s KILL 200 python bench.py --no-cpu-baseline --no-f32-line --roi-layout $lay > gpurun_out/b_$lay.json 2> gpurun_out/b_$lay.err; echo "rc=$?"
done
python - <<'P'
import json
for n in ("rchw","rhwc"):
    try:
        d=json.loads(open(f"gpurun_out/b_{n}.json").read().strip().splitlines()[-1])
        k=d["kernels"]
        print(n, d["ms_per_step"], d["e2e"]["ms_per_step"], {x:(k[x]["ms"],k[x]["frac"]) for x in k if x.startswith("roi_align")})
    except Exception as e:
        print(n, "failed", e)
P
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
tail -4 gpurun_out/t_all.log
