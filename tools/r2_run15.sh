#!/bin/bash
# source-level ncu capture of the bounce-1 kernels of config 2 (where do the instructions go?)
python bench.py --one-step --scene cornell_monkey > gpurun_out/onestep_c2.log 2>&1 || exit 1
PTB_NO_OVERLAP=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:^(k_trace_pre|k_trace_tree|k_shade)" -s 5 -c 5 -o gpurun_out/prof_r2b_c2 -f python bench.py --one-step --scene cornell_monkey > gpurun_out/ncu_c2.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/
