#!/bin/bash
PTB_NO_RESIDENT_BVH=1 python -m pytest tests -q -m gpu > gpurun_out/tests_global.log 2>&1; tail -4 gpurun_out/tests_global.log
python -m pytest tests -q -m gpu > gpurun_out/tests_default.log 2>&1; tail -3 gpurun_out/tests_default.log
python tools/mega_sweep.py mega 32 2>&1 | tail -5
