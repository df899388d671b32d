#!/bin/bash
# Round-2 call 2 (N GPUs, N = 2 first, then 8):   gpurun --gpus N --timeout 1200 -- 'bash scripts/r02_call2_multi_gpu.sh N'
# NCCL/IPC legs of the halo modes (bit-identity tests), then the partitioned C5 step per halo mode on the uniform and the
# locality graph, then the data-parallel C4 step.  One file per run under gpurun_out/; failures do not stop the script.
set -u
N=${1:-2}
MODE=${2:-full}      # "quick" = the five most informative runs (use it at 8 GPUs: every minute there costs 8 GPU-minutes)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { local name=$1; shift; echo "== $name"; ( timeout 300 "$@" ) > "gpurun_out/r02_n${N}_$name.log" 2>&1; echo "rc=$? n$N $name" | tee -a gpurun_out/r02_call2_status.txt; }

[ "$N" = 2 ] && GNNB200_RUN_UNVERIFIED=1 run partition_tests python -m pytest tests/test_gpu_partition.py -m gpu -q --tb=short -p no:cacheprovider
port=29600
LOC="dense sparse peer peercopy"; [ "$MODE" = quick ] && LOC="sparse peer"
for h in dense peercopy; do port=$((port+1)); run c5_uniform_$h $T --master-port $port bench.py --gpus $N --halo $h --steps 5 --warmup 3; done
for h in $LOC; do port=$((port+1)); run c5_loc09_$h $T --master-port $port bench.py --gpus $N --locality 0.9 --halo $h --steps 5 --warmup 3; done
port=$((port+1)); run c4 $T --master-port $port bench.py --gpus $N --workload c4 --steps 20 --warmup 5
cat gpurun_out/r02_call2_status.txt
