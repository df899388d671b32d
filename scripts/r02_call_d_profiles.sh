#!/bin/bash
# Round-2 call D (ONE GPU, ncu only): launch list of the DEFAULT bench command at full scale, and `--set full` captures of
# the kernel families of one training step at quarter scale (612 k rows: far beyond L2) — the encoder + first layer's forward,
# then the last layer's backward.  The reports are exported to raw CSV on the box and deleted (gpurun_out is capped at 64 MiB).
set -u
mkdir -p gpurun_out
FULL="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e --no-selfcheck --no-generator2"
QUARTER="$FULL --scale 0.25"
$FULL > gpurun_out/r02d_plain_full.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02d_launches_full_scale.csv $FULL > gpurun_out/r02d_ncu_launches.log 2>&1
echo "launch list rc=$?"
K='regex:gemm_tf32_kernel|bn_act|colstats_partial|aggregate_vec|splitk_reduce'
$QUARTER > gpurun_out/r02d_plain_quarter.log 2>&1 &&
ncu --set full --clock-control none -k "$K" -c 14 -o /tmp/r02d_fwd -f $QUARTER > gpurun_out/r02d_ncu_fwd.log 2>&1
echo "fwd capture rc=$?"
ncu -i /tmp/r02d_fwd.ncu-rep --page raw --csv > gpurun_out/r02d_fwd_raw.csv 2>/dev/null
ncu --set full --clock-control none -k "$K" -s 38 -c 18 -o /tmp/r02d_bwd -f $QUARTER > gpurun_out/r02d_ncu_bwd.log 2>&1
echo "bwd capture rc=$?"
ncu -i /tmp/r02d_bwd.ncu-rep --page raw --csv > gpurun_out/r02d_bwd_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out | grep r02d
