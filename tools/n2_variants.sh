#!/bin/bash
# 2-GPU bench under a few NCCL settings (tools only)
run() {
  echo "== $1"
  env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | grep -E "^\{|NCCL INFO.*(Algo|algo|NVLS|channels|Channel 00/)" | cut -c1-170 | tail -4
}
run "X=1"
run "NCCL_ALGO=NVLS"
run "NCCL_MAX_NCHANNELS=8"
run "NCCL_MAX_NCHANNELS=4"
run "NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL"
