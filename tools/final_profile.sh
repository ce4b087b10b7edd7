#!/bin/bash
# ncu evidence for the final build of a round (run under gpurun, after bench.py has exited 0 without ncu)
python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:^(k_trace_pre|k_trace_tree|k_shade)" -s 30 -c 5 -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/launches_final.csv gpurun_out/prof_final.ncu-rep
