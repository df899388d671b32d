#!/bin/bash
# Round-2 call L (ONE GPU): validation of the committed state + experiment: BatchNorm statistics from the GEMM epilogue.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02l_status.txt
run() { local name=$1; shift; echo "== $name" ; ( timeout "${LIMIT:-300}" "$@" ) > "gpurun_out/r02l_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02l_status.txt; }
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary --no-generator2 --no-e2e"
GNNB200_NATIVE_LAYER=0 run c5_python_layers $B
GNNB200_NATIVE_LAYER=0 GNNB200_GEMM_STATS=1 run c5_python_layers_gemm_stats $B
GNNB200_NATIVE_LAYER=0 GNNB200_GEMM_STATS=1 run tests_gemm_stats python -m pytest -m gpu -q --tb=short -p no:cacheprovider tests/test_gpu_models.py tests/test_gpu_fused.py
LIMIT=900 run suite python -m pytest -m gpu -q -x --tb=short -p no:cacheprovider tests
LIMIT=600 run c5_full python bench.py --steps 5 --warmup 3
cat gpurun_out/r02l_status.txt
