#!/bin/bash
N=${1:-2}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --grad-sync peer --peer-ctas 64 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])"; }
echo "base $(run)"
echo "noflat $(DA_DIAG_NOFLAT=1 run)"
echo "nolimit $(DA_DIAG_NOLIMIT=1 run)"
echo "nopeer $(DA_DIAG_NOPEER=1 run)"
echo "nopeer+noflat $(DA_DIAG_NOPEER=1 DA_DIAG_NOFLAT=1 run)"
