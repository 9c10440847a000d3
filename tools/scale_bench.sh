#!/bin/bash
# tools/scale_bench.sh <N> [extra bench.py flags]: peer_check + bench.py on N GPUs of one box (gpurun --gpus N -- 'tools/scale_bench.sh N').
# Writes gpurun_out/n<N>_bench.json and prints ms/step, pairs/s, e2e pairs/s.
N=${1:-2}; shift
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/peer_check.py 2>&1 | grep "PEER_\|Error" | head -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N "$@" 2>gpurun_out/n${N}_err.log | grep '^{' > gpurun_out/n${N}_bench.json
python -c "import json; d=json.load(open('gpurun_out/n${N}_bench.json')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['grad_allreduce'][:80])"
grep -i "error\|Traceback" gpurun_out/n${N}_err.log | head -5
