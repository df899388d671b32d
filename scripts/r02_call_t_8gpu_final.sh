#!/bin/bash
# Round-2 call T (EIGHT GPUs): the driver's N = 8 command on the final tree.
set -u
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
( timeout 300 $T --master-port 29841 bench.py --gpus 8 --steps 5 --warmup 3 ) > gpurun_out/r02t_c5_default_n8.log 2>&1; echo "rc=$? c5 n8"
tail -1 gpurun_out/r02t_c5_default_n8.log | cut -c1-400
