#!/bin/bash
# Round-2 call O (ONE GPU): is a 1/8-size C5 step (what one rank of an 8-GPU run computes) host-bound?  Same step through the
# C++ layer composite, the Python one-node-per-layer path and the one-node-per-kernel path the node partition uses.
set -u
mkdir -p gpurun_out
B="--scale 0.125 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary --no-generator2 --no-e2e --no-selfcheck"
python bench.py $B > gpurun_out/r02o_native.log 2>&1; echo "rc=$? native"
GNNB200_NATIVE_LAYER=0 python bench.py $B > gpurun_out/r02o_python_fused.log 2>&1; echo "rc=$? python fused"
python -c "
import sys, runpy
sys.path.insert(0, '.')
import gnnb200
from gnnb200 import models
models.GINLayer.fused = False
sys.argv = ['bench.py'] + '$B'.split()
runpy.run_path('bench.py', run_name='__main__')
" > gpurun_out/r02o_op_level.log 2>&1; echo "rc=$? op level"
for f in native python_fused op_level; do tail -1 gpurun_out/r02o_$f.log | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('$f', d['ms_per_step'], d['gpu_launches'] / d['steps'])"; done
