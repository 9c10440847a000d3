#!/bin/bash
N=${1:-4}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/peer_check.py --time 2>&1 | grep "PEER\|NCCL_TIME\|Error\|error"
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" 2>gpurun_out/n${N}_err.log | grep '^{' ; }
run --grad-sync peer > gpurun_out/n${N}_peer_copy.json
run --grad-sync nccl > gpurun_out/n${N}_nccl.json
for f in gpurun_out/n${N}_peer_copy.json gpurun_out/n${N}_nccl.json; do echo "$f $(python -c "import json,sys; d=json.load(open('$f')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['grad_allreduce'][:70])")"; done
