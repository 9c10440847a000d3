#!/bin/bash
# N-GPU bench variants: tools/n2_bench.sh <N>  (results in gpurun_out/n<N>_*.json)
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" 2>gpurun_out/n${N}_err.log | grep '^{' ; }
run --grad-sync peer > gpurun_out/n${N}_peer_copy.json
run --grad-sync peer --peer-transport stores --peer-ctas 64 --peer-reserve-sms 0 > gpurun_out/n${N}_peer_stores.json
run --grad-sync nccl > gpurun_out/n${N}_nccl.json
for f in gpurun_out/n${N}_*.json; do echo "$f $(python -c "import json,sys; d=json.load(open('$f')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['grad_allreduce'][:70])")"; done
tail -5 gpurun_out/n${N}_err.log
