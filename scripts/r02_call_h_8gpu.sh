#!/bin/bash
# Round-2 call H (EIGHT GPUs): the DEFAULT bench command at N = 8 and N = 4 (halo 'auto': uniform graph -> all-gather
# pipelined over column slabs; generator (ii) -> overlapped sparse exchange), what the driver's scaling run will execute.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02h_status.txt
run() { local name=$1; shift; echo "== $name"; ( timeout 300 "$@" ) > "gpurun_out/r02h_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02h_status.txt; }
T() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2"; }
run c5_default_n8 $(T 8 29901) bench.py --gpus 8 --steps 5 --warmup 3
run c5_default_n4 $(T 4 29902) bench.py --gpus 4 --steps 5 --warmup 3
GNNB200_HALO_SLABS=4 run c5_slabs4_n8 $(T 8 29903) bench.py --gpus 8 --halo dense --steps 3 --warmup 2 --no-generator2 --no-selfcheck --no-e2e
cat gpurun_out/r02h_status.txt
