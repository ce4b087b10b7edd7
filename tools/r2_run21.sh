#!/bin/bash
# the driver's own calls: default bench (N=1), reference arm, smoke
t0=$(date +%s)
python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo "bench exit $? in $(( $(date +%s) - t0 )) s"; tail -c 300 gpurun_out/bench_r2_final.err
t0=$(date +%s)
python bench.py --impl reference > gpurun_out/bench_r2_reference.json 2>/dev/null; echo "reference exit $? in $(( $(date +%s) - t0 )) s"; tail -c 500 gpurun_out/bench_r2_reference.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
