#!/bin/bash
# Round-2 call J (ONE GPU): the pre-split-weight GEMM with 16-wide K steps (four-stage ring) against the 32-wide two-stage ring.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02j_status.txt
run() { local name=$1; shift; echo "== $name" ; ( timeout "${LIMIT:-300}" "$@" ) > "gpurun_out/r02j_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02j_status.txt; }
GNNB200_X3W_BK=16 run tests_bk16 python -m pytest -m gpu -q --tb=short -p no:cacheprovider tests/test_gpu_gemm.py tests/test_gpu_models.py tests/test_gpu_fused.py
GNNB200_X3W_BK=16 run gemm_bk16 python scripts/bench_gemm.py
run gemm_bk32 python scripts/bench_gemm.py
GNNB200_X3W_BK=16 run c5_bk16 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-generator2
cat gpurun_out/r02j_status.txt
