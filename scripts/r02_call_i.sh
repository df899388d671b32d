#!/bin/bash
# Round-2 call I (ONE GPU): final validation of the committed state (suite, smoke, default bench line) + one GEMM experiment.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02i_status.txt
run() { local name=$1; shift; echo "== $name" ; ( timeout "${LIMIT:-300}" "$@" ) > "gpurun_out/r02i_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02i_status.txt; }
LIMIT=900 run suite python -m pytest -m gpu -q -x --tb=short -p no:cacheprovider tests
run smoke python __graft_entry__.py smoke
run gemm_bn256 python scripts/bench_gemm.py
GNNB200_X3W_BN=128 run gemm_bn128 python scripts/bench_gemm.py
LIMIT=600 run c5_full python bench.py --steps 5 --warmup 3
run c1 python bench.py --workload c1
cat gpurun_out/r02i_status.txt
