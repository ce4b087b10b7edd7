#!/bin/bash
python -m pytest tests -q -m gpu > gpurun_out/tests_default.log 2>&1; tail -6 gpurun_out/tests_default.log
python tools/mega_sweep.py mega 32 2>&1 | tail -12
python tools/mega_sweep.py matball 32 2>&1 | tail -3
