#!/bin/bash
python -m pytest tests -q -m gpu > gpurun_out/tests_default.log 2>&1; tail -6 gpurun_out/tests_default.log
python tools/mega_sweep.py mega 32 2>&1 | tail -13
for k in 1 2 3 4; do python tools/mlt_bench.py 64 $k 2>&1 | tail -1; done
