#!/bin/bash
# usage: mgpu_sweep.sh NGPUS "ENV1=.. ENV2=.." "..." : ms/step of bench.py at 16384^2 for each environment
n=$1; shift
for spec in "$@"; do
  env $spec python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 8 --warmup 3 --cpu-steps 0 --eager-steps 0 --secondary-n 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('%-60s %.3f ms  xor %s' % ('$spec', d['ms_per_step'], d['fingerprint']['omega_hat_xor']))"
done
