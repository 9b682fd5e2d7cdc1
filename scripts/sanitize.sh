#!/bin/bash
# compute-sanitizer runs of the smoke scene: memcheck on every kernel, racecheck on the shared-memory
# traversal stacks.  Run on the GPU box:  bash scripts/sanitize.sh  -> gpurun_out/sanitize_{memcheck,racecheck}.txt
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
PY='import sys; sys.path.insert(0, "."); import __graft_entry__ as g; g.smoke(); import scripts.sanitize_extra as x; x.run()'
for tool in memcheck racecheck; do
    timeout 900 $SAN --tool $tool --print-limit 20 python -c "$PY" > gpurun_out/sanitize_$tool.txt 2>&1
    echo "$tool rc=$?" >> gpurun_out/sanitize_$tool.txt
    tail -4 gpurun_out/sanitize_$tool.txt
done
