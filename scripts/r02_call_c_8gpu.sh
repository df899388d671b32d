#!/bin/bash
# Round-2 call C (EIGHT GPUs; every minute costs 8 GPU-minutes): the node-partitioned C5 step per halo mode on both generators
# at N = 8, and the data-parallel C4 step at N = 1, 2, 4, 8 on the same box (its weak-scaling curve).
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02c_status.txt
run() { local name=$1; shift; echo "== $name"; ( timeout 240 "$@" ) > "gpurun_out/r02c_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02c_status.txt; }
T() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2"; }
port=29700
for n in 1 2 4 8; do port=$((port+1)); run c4_n$n $(T $n $port) bench.py --gpus $n --workload c4 --steps 20 --warmup 5; done
for h in dense peercopy; do port=$((port+1)); run c5_uniform_n8_$h $(T 8 $port) bench.py --gpus 8 --halo $h --steps 5 --warmup 3; done
for h in sparse peer; do port=$((port+1)); run c5_loc09_n8_$h $(T 8 $port) bench.py --gpus 8 --locality 0.9 --halo $h --steps 5 --warmup 3; done
port=$((port+1)); run c5_loc09_n4_peer $(T 4 $port) bench.py --gpus 4 --locality 0.9 --halo peer --steps 5 --warmup 3
cat gpurun_out/r02c_status.txt
