#!/bin/bash
# Round-2 call M (ONE GPU): what the driver runs at round end, on the committed state: GPU suite (-x), smoke, default bench, reference arm.
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02m_status.txt
run() { local name=$1; shift; echo "== $name" ; ( timeout "${LIMIT:-300}" "$@" ) > "gpurun_out/r02m_$name.log" 2>&1; echo "rc=$? $name" | tee -a gpurun_out/r02m_status.txt; }
LIMIT=900 run suite python -m pytest tests/ -x -q -m gpu -p no:cacheprovider
run smoke python -c "import __graft_entry__ as g; g.smoke()"
LIMIT=600 run bench python bench.py
cat gpurun_out/r02m_status.txt
