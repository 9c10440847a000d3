#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout -s KILL 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout -s KILL 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "rc=$?" >> gpurun_out/t_all.log
tail -3 gpurun_out/t_all.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/smoke.log
python - <<'P'
import json
d=json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
k=d["kernels"]
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["clocks"])
print({x:(k[x]["ms"],k[x]["frac"]) for x in k if x.startswith("roi_align")})
P
