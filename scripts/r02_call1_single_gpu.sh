#!/bin/bash
# Round-2 call 1 (ONE GPU, ~6 min of box time): everything that was written without GPU access, one variant at a time,
# each into its own file under gpurun_out/.   gpurun --timeout 1200 -- 'bash scripts/r02_call1_single_gpu.sh'
# A variant that fails or hangs must not take the others down: every step has its own timeout and `|| true`.
set -u
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider"
run() { local name=$1; shift; echo "== $name" ; ( timeout "${LIMIT:-300}" "$@" ) > "gpurun_out/r02_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02_call1_status.txt; }

run suite_default            $P tests
LIMIT=1200 GNNB200_RUN_UNVERIFIED=1 run suite_unverified $P tests     # includes the every-row C5 check against the C oracle (~2 min of host time)
GNNB200_EW_V2=1              run suite_ew_v2              $P tests/test_gpu_bn.py tests/test_gpu_models.py tests/test_gpu_fused.py tests/test_gpu_elementwise_v2.py
GNNB200_GEMM_TMA_STORE=1     run suite_tma_store          $P tests/test_gpu_gemm.py tests/test_gpu_models.py tests/test_gpu_fused.py
GNNB200_NATIVE_LAYER=1       run suite_native_layer       $P tests/test_gpu_fused.py tests/test_gpu_models.py tests/test_gpu_finetune_step.py tests/test_gpu_pretrain_step.py
run elementwise_v1_v2        python scripts/bench_elementwise.py

B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
run bench_c5_default         $B --no-secondary
GNNB200_EW_V2=1              run bench_c5_ew_v2           $B --no-secondary
GNNB200_GEMM_TMA_STORE=1     run bench_c5_tma_store       $B --no-secondary
GNNB200_EW_V2=1 GNNB200_GEMM_TMA_STORE=1 run bench_c5_both $B --no-secondary
run bench_secondary_default  $B
GNNB200_NATIVE_LAYER=1       run bench_secondary_native   $B
run bench_c4_n1              python bench.py --workload c4 --steps 20 --warmup 5
GNNB200_NATIVE_LAYER=1       run bench_c4_n1_native       python bench.py --workload c4 --steps 20 --warmup 5
run bench_c5_locality09      $B --no-secondary --locality 0.9
run bench_c5_skew18          $B --no-secondary --locality 0.9 --skew 1.8
GNNB200_LONG_ROWS=1          run bench_c5_skew18_longrows $B --no-secondary --locality 0.9 --skew 1.8
cat gpurun_out/r02_call1_status.txt
