#!/bin/bash
# Evidence of the final round-2 build (run under gpurun; every ncu pass follows a plain run of the same command that exited 0):
#   1. GPU tests in the three node modes            2. L2 probe, plain and under ncu
#   3. instruction counts of one step per config (-> profiles/inst_counts.json via tools/ncu_inst_counts.py)
#   4. launch lists (config 2, config 4)             5. --set full captures: config 2 bounce-1 kernels, config 4 tree kernels
#   6. the driver's calls: bench.py, bench.py --impl reference, smoke()
python -m pytest tests -q -m gpu > gpurun_out/tests_default.log 2>&1; tail -1 gpurun_out/tests_default.log
PTB_QUANT_RESIDENT_BVH=1 python -m pytest tests -q -m gpu > gpurun_out/tests_qres.log 2>&1; tail -1 gpurun_out/tests_qres.log
PTB_NO_RESIDENT_BVH=1 python -m pytest tests -q -m gpu > gpurun_out/tests_global.log 2>&1; tail -1 gpurun_out/tests_global.log
python tools/l2_probe.py > gpurun_out/l2_probe.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,l1tex__t_bytes.sum -k regex:k_l2_read --clock-control none --csv --log-file gpurun_out/l2_probe_ncu.csv python tools/l2_probe.py > /dev/null 2>&1
cat gpurun_out/l2_probe.log
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for sc in cornell_monkey cornell_boxes matball; do
  python bench.py --one-step --scene $sc > gpurun_out/onestep_$sc.log 2>&1 && \
  ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/inst_$sc.csv python bench.py --one-step --scene $sc > gpurun_out/onestep_ncu_$sc.log 2>&1
  echo "$sc inst ncu exit $?"
done
python bench.py --one-step --scene mega --spp 4 > gpurun_out/onestep_mega.log 2>&1 && \
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/inst_mega.csv python bench.py --one-step --scene mega --spp 4 > gpurun_out/onestep_ncu_mega.log 2>&1
echo "mega inst ncu exit $?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_c2.csv python bench.py --one-step --scene cornell_monkey > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_mega.csv python bench.py --one-step --scene mega --spp 4 > /dev/null 2>&1
PTB_NO_OVERLAP=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:^(k_trace_pre|k_trace_tree|k_shade)" -s 5 -c 5 -o gpurun_out/prof_r2_final_c2 -f python bench.py --one-step --scene cornell_monkey > gpurun_out/ncu_c2.log 2>&1
echo "c2 full ncu exit $?"
PTB_NO_OVERLAP=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:^k_trace_tree" -c 4 -o gpurun_out/prof_r2_final_mega -f python bench.py --one-step --scene mega --spp 4 > gpurun_out/ncu_mega.log 2>&1
echo "mega full ncu exit $?"
bash tools/r2_run21.sh
