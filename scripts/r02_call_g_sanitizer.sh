#!/bin/bash
# Round-2 call G (ONE GPU, compute-sanitizer memcheck only — one tool per call, B200_PROFILING.md): the kernels written in
# round 2 (pre-split-weight GEMM incl. its mbarrier / TMEM pipeline, head tails and loss sums, negative sampling, BatchNorm
# merge, hub rows) on their smallest test cases.
set -u
mkdir -p gpurun_out
SEL='tests/test_gpu_heads.py tests/test_gpu_negative_sampling.py tests/test_gpu_pool.py tests/test_gpu_structure.py'
K='not 70001 and not 20000 and not 4100 and not 4000 and not 300000 and not 600000 and not 127000 and not 2100 and not 724'
python -m pytest $SEL tests/test_gpu_gemm.py tests/test_gpu_bn.py tests/test_gpu_aggregate.py -m gpu -q -x -p no:cacheprovider -k "$K and not 2708 and not 3327 and not 9000 and not 60000 and not 50000 and not 5000 and not 4096" > gpurun_out/r02g_plain.log 2>&1
echo "plain rc=$?" | tee gpurun_out/r02g_status.txt
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
  python -m pytest $SEL tests/test_gpu_gemm.py tests/test_gpu_bn.py tests/test_gpu_aggregate.py -m gpu -q -x -p no:cacheprovider \
  -k "$K and not 2708 and not 3327 and not 9000 and not 60000 and not 50000 and not 5000 and not 4096" > gpurun_out/r02g_memcheck.log 2>&1
echo "memcheck rc=$?" | tee -a gpurun_out/r02g_status.txt
tail -5 gpurun_out/r02g_memcheck.log
