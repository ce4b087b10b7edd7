#!/bin/bash
# ncu call: instruction counts of one step for configs 1-3, full capture of the config-4 tree kernels and of config 2's top kernels
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for sc in cornell_monkey cornell_boxes matball; do
  python bench.py --one-step --scene $sc > gpurun_out/onestep_$sc.log 2>&1 && \
  ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/inst_$sc.csv python bench.py --one-step --scene $sc > gpurun_out/onestep_ncu_$sc.log 2>&1
  echo "$sc ncu exit $?"
done
python bench.py --one-step --scene mega --spp 4 > gpurun_out/onestep_mega.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:^k_trace_tree" -c 4 -o gpurun_out/prof_r2_mega -f python bench.py --one-step --scene mega --spp 4 > gpurun_out/ncu_mega.log 2>&1
echo "mega ncu exit $?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:^(k_trace_pre|k_trace_tree|void .*k_shade)" -s 3 -c 5 -o gpurun_out/prof_r2_c2 -f python bench.py --one-step --scene cornell_monkey > gpurun_out/ncu_c2.log 2>&1
echo "c2 ncu exit $?"
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_mega.csv python bench.py --one-step --scene mega --spp 4 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/inst_*.csv | tail
