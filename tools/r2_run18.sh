#!/bin/bash
# shade kernel: sorted vs queue order, instruction counts and stalls (ncu, bounce-1 launch)
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,dram__bytes_read.sum
for v in default nosort; do
  lib=""; [ $v != default ] && lib="PTINA_B200_LIB=$PWD/variants/$v.so"
  env $lib PTB_NO_OVERLAP=1 ncu --profile-from-start off --metrics $M --clock-control none -k "regex:k_shade" --csv --log-file gpurun_out/shade_$v.csv python bench.py --one-step --scene cornell_monkey > /dev/null 2>&1
  echo "== $v"; python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/shade_$v.csv')) if len(r)>10]
h=rows[0]; ix={k:i for i,k in enumerate(h)}
by={}
for r in rows[1:]:
    by.setdefault(r[ix['ID']],{})[r[ix['Metric Name']]]=r[ix['Metric Value']]
for i,(k,m) in enumerate(by.items()):
    print(i, {a.split('.')[0].replace('smsp__average_warps_issue_stalled_','st_').replace('_per_issue_active',''):b for a,b in m.items()})
PY
done
