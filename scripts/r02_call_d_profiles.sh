#!/bin/bash
# Round-2 call D (ONE GPU, ncu only): launch list of the DEFAULT bench command at full scale, and `--set full` captures of
# every kernel family of one training step at quarter scale (612 k rows: far beyond L2) — forward pass, then backward pass.
set -u
mkdir -p gpurun_out
FULL="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e --no-selfcheck --no-generator2"
QUARTER="$FULL --scale 0.25"
$FULL > gpurun_out/r02d_plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02d_launches_full_scale.csv $FULL > gpurun_out/r02d_ncu_launches.log 2>&1
echo "launch list rc=$?"
K='regex:gemm_tf32_kernel|bn_act|colstats_partial|aggregate_vec|splitk_reduce|dot_partial'
$QUARTER > gpurun_out/r02d_plain_quarter.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "$K" -c 40 -o gpurun_out/r02d_fwd -f $QUARTER > gpurun_out/r02d_ncu_fwd.log 2>&1
echo "fwd capture rc=$?"
ncu --set full --clock-control none --import-source on -k "$K" -s 40 -c 60 -o gpurun_out/r02d_bwd -f $QUARTER > gpurun_out/r02d_ncu_bwd.log 2>&1
echo "bwd capture rc=$?"
ls -la gpurun_out | grep r02d
