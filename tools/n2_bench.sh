#!/bin/bash
# N-GPU bench variants: tools/n2_bench.sh <N>  (results in gpurun_out/n<N>_*.json)
N=${1:-2}
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" 2>gpurun_out/n${N}_err.log | grep '^{' ; }
run --grad-sync peer > gpurun_out/n${N}_peer_deferred.json
run --grad-sync peer --peer-publish in-step > gpurun_out/n${N}_peer_instep.json
for f in gpurun_out/n${N}_peer_deferred.json gpurun_out/n${N}_peer_instep.json; do echo "$f $(python -c "import json,sys; d=json.load(open('$f')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['grad_allreduce'][:50])")"; done
grep -i "error\|Traceback" gpurun_out/n${N}_err.log | head -5
