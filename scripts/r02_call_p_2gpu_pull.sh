#!/bin/bash
# Round-2 call P (TWO GPUs): 'sparse_pull' over CUDA IPC / NVLink: bit-identity test, then the step against 'sparse_overlap'.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02p_status.txt
run() { local name=$1; shift; echo "== $name"; ( timeout 300 "$@" ) > "gpurun_out/r02p_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02p_status.txt; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run partition_tests python -m pytest tests/test_gpu_partition.py -m gpu -q --tb=short -p no:cacheprovider
run c5_loc09_pull $T --master-port 29821 bench.py --gpus 2 --locality 0.9 --skew 1.8 --halo sparse_pull --steps 5 --warmup 3 --no-e2e
run c5_loc09_overlap $T --master-port 29822 bench.py --gpus 2 --locality 0.9 --skew 1.8 --halo sparse_overlap --steps 5 --warmup 3 --no-e2e
cat gpurun_out/r02p_status.txt
