#!/bin/bash
N=${1:-8}
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/peer_check.py 2>&1 | grep "PEER\|Error\|error" | head -5
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N 2>gpurun_out/n${N}_err.log | grep '^{' > gpurun_out/n${N}_peer_copy.json
python -c "import json,sys; d=json.load(open('gpurun_out/n${N}_peer_copy.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['grad_allreduce'][:70])"
tail -3 gpurun_out/n${N}_err.log
