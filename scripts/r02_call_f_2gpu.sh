#!/bin/bash
# Round-2 call F (TWO GPUs): the partition tests with the column-slab and overlapped-sparse exchanges, then the default
# bench command at N = 2 (halo 'auto': uniform graph -> dense; generator (ii) -> sparse_overlap) and explicit variants.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02f_status.txt
run() { local name=$1; shift; echo "== $name"; ( timeout 400 "$@" ) > "gpurun_out/r02f_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02f_status.txt; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run partition_tests python -m pytest tests/test_gpu_partition.py -m gpu -q --tb=short -p no:cacheprovider
run c5_default $T --master-port 29801 bench.py --gpus 2 --steps 5 --warmup 3
run c5_loc09_sparse_overlap $T --master-port 29803 bench.py --gpus 2 --locality 0.9 --halo sparse_overlap --steps 5 --warmup 3 --no-generator2
GNNB200_HALO_SLABS=2 run c5_dense_2slabs $T --master-port 29804 bench.py --gpus 2 --halo dense --steps 5 --warmup 3 --no-generator2 --no-e2e
cat gpurun_out/r02f_status.txt
