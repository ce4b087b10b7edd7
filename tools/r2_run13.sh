#!/bin/bash
# final evidence call: L2 probe (plain, then ncu), launch list and instruction counts of the final build, full bench
python tools/l2_probe.py > gpurun_out/l2_probe.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,lts__t_bytes.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,l1tex__t_bytes.sum -k regex:k_l2_read --clock-control none --csv --log-file gpurun_out/l2_probe_ncu.csv python tools/l2_probe.py > /dev/null 2>&1
cat gpurun_out/l2_probe.log
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for sc in cornell_monkey cornell_boxes matball; do
  python bench.py --one-step --scene $sc > gpurun_out/onestep_$sc.log 2>&1 && \
  ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/inst_$sc.csv python bench.py --one-step --scene $sc > gpurun_out/onestep_ncu_$sc.log 2>&1
  echo "$sc ncu exit $?"
done
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2_c2.csv python bench.py --one-step --scene cornell_monkey > /dev/null 2>&1
python bench.py --one-step --scene mega --spp 4 > gpurun_out/onestep_mega.log 2>&1 && \
ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:^k_trace_tree" -c 4 -o gpurun_out/prof_r2_mega_w4 -f python bench.py --one-step --scene mega --spp 4 > gpurun_out/ncu_mega.log 2>&1
echo "mega ncu exit $?"
python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo "bench exit $?"; tail -c 400 gpurun_out/bench_r2_final.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_reference.json 2>/dev/null; tail -c 600 gpurun_out/bench_r2_reference.json
