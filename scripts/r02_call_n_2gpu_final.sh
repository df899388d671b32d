#!/bin/bash
# Round-2 call N (TWO GPUs): the committed state through the 2-GPU tests and the default multi-GPU bench command.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02n_status.txt
run() { local name=$1; shift; echo "== $name"; ( timeout 400 "$@" ) > "gpurun_out/r02n_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02n_status.txt; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run partition_tests python -m pytest tests/test_gpu_partition.py -m gpu -q --tb=short -p no:cacheprovider
run c5_default $T --master-port 29811 bench.py --gpus 2 --steps 5 --warmup 3
run c4 $T --master-port 29812 bench.py --gpus 2 --workload c4 --steps 10 --warmup 3
run reference_arm $T --master-port 29813 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 --cpu-sample 0.03
cat gpurun_out/r02n_status.txt
