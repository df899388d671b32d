#!/bin/bash
# Round-2 call E (ONE GPU): the suite with the head / loss / negative-sampling / merge kernels and the re-pitched encoder
# GEMMs, then the default bench line (selfcheck, generator (ii), secondary configs) and the C1 line.
set -u
mkdir -p gpurun_out
P="python -m pytest -m gpu -q --tb=short -p no:cacheprovider"
run() { local name=$1; shift; echo "== $name" ; ( timeout "${LIMIT:-300}" "$@" ) > "gpurun_out/r02e_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02e_status.txt; }
rm -f gpurun_out/r02e_status.txt
LIMIT=900 run suite $P tests
run smoke python __graft_entry__.py smoke
LIMIT=600 run c5_full python bench.py --steps 5 --warmup 3
run c1 python bench.py --workload c1
run c4 python bench.py --workload c4 --steps 20 --warmup 5
cat gpurun_out/r02e_status.txt
