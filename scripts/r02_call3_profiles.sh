#!/bin/bash
# Round-2 call 3 (ONE GPU): the profiles the judge asks for, taken from the DEFAULT bench command at full scale.
#   gpurun --timeout 1500 -- 'bash scripts/r02_call3_profiles.sh'     then copy the summaries into profiles/r02_*.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --no-e2e"
$CMD > gpurun_out/r02_plain_before_ncu.log 2>&1 || exit 1          # a number under ncu is never a bench value
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_full_scale.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
for k in aggregate_vec gemm_tf32 bn_act_bwd_apply bn_act_fwd colstats_partial; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/r02_$k -f $CMD > gpurun_out/r02_ncu_$k.log 2>&1
done
# small-graph steps: sum of kernel times vs wall time per step tells whether C2/C3 are host- or device-bound
python bench.py --only-secondary --no-cpu-baseline > gpurun_out/r02_secondary_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r02_launches_secondary.csv python bench.py --only-secondary --no-cpu-baseline --secondary-steps 2 > gpurun_out/r02_ncu_secondary.log 2>&1
ls -la gpurun_out | tail -20
