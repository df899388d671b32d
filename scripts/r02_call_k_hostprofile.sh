#!/bin/bash
# Round-2 call K (ONE GPU): where the HOST time of the small-graph steps goes (cProfile of the C4 step on a real device).
set -u
mkdir -p gpurun_out
python -c "
import cProfile, pstats, sys, io, torch
sys.argv = ['bench.py']
import importlib.util
spec = importlib.util.spec_from_file_location('bench', 'bench.py'); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
dev = torch.device('cuda', 0)
step, sampler = b.build_c4_step(dev, 0, 1)
fixed = sampler.draw()
for _ in range(5): step(fixed)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(10): step(fixed)
torch.cuda.synchronize()
print('plain ms/step', (time.perf_counter() - t0) * 100)
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step(fixed)
torch.cuda.synchronize()
pr.disable()
for key in ('tottime', 'cumtime'):
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats(key).print_stats(45); print(s.getvalue()[:9000])
" > gpurun_out/r02k_c4_cprofile.txt 2>&1
echo "rc=$?"
tail -3 gpurun_out/r02k_c4_cprofile.txt
