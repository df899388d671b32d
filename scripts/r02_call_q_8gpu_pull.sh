#!/bin/bash
# Round-2 call Q (EIGHT GPUs): generator (ii) with the halo rows pulled over NVLink ('sparse_pull') against the all-to-all.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02q_status.txt
run() { local name=$1; shift; echo "== $name"; ( timeout 240 "$@" ) > "gpurun_out/r02q_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02q_status.txt; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
B="bench.py --gpus 8 --locality 0.9 --skew 1.8 --steps 5 --warmup 3 --no-e2e"
run pull $T --master-port 29831 $B --halo sparse_pull
run overlap $T --master-port 29832 $B --halo sparse_overlap
cat gpurun_out/r02q_status.txt
