#!/bin/bash
# Round-2 call B (ONE GPU): the suite with the new defaults (column-stationary BN kernels, TMA-store epilogue, hub rows, layer
# composites, forward GEMMs 3xTF32 on pre-split weights), GEMM microbenchmark, C5 per precision, small-graph configs.
set -u
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider"
run() { local name=$1; shift; echo "== $name" ; ( timeout "${LIMIT:-300}" "$@" ) > "gpurun_out/r02b_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02b_status.txt; }
rm -f gpurun_out/r02b_status.txt
LIMIT=900 run suite $P tests
run smoke python __graft_entry__.py smoke
run gemm python scripts/bench_gemm.py
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
run c5_fwd3 $B
GNNB200_X3W_RAW_HI=1 run c5_fwd3_rawhi $B
run c5_tf32 $B --precision tf32
run c5_skew18 $B --locality 0.9 --skew 1.8
LIMIT=600 run c5_full python bench.py --steps 5 --warmup 3
run c1 python bench.py --workload c1
LIMIT=400 run reference_arm python bench.py --impl reference --steps 3 --warmup 1
cat gpurun_out/r02b_status.txt
