#!/bin/bash
summ='import json,sys
d=json.loads(sys.stdin.read())
print("%-12s Mrays/s %7.1f  ms/step %6.3f  e2e %7.1f  stages %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}))'
python -m pytest tests -q -m gpu > gpurun_out/tests_default.log 2>&1; tail -4 gpurun_out/tests_default.log
PTB_NO_RESIDENT_BVH=1 python -m pytest tests -q -m gpu > gpurun_out/tests_global.log 2>&1; tail -3 gpurun_out/tests_global.log
PTB_QUANT_RESIDENT_BVH=1 python -m pytest tests -q -m gpu > gpurun_out/tests_quant.log 2>&1; tail -3 gpurun_out/tests_quant.log
for v in fmadonly fastdivonly; do
  echo "== $v"
  PTINA_B200_LIB=$PWD/variants/$v.so python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_fullsize.py -q -m gpu > gpurun_out/tests_$v.log 2>&1; tail -12 gpurun_out/tests_$v.log
  PTINA_B200_LIB=$PWD/variants/$v.so python bench.py --quick --no-cpu 2>gpurun_out/ab_$v.err | tail -1 | python -c "$summ" $v
done
echo "== mega stream A/B"
python bench.py --quick --no-cpu --scene mega --steps 3 --warmup 1 2>gpurun_out/ab_mega.err | tail -1 | python -c "$summ" default
PTINA_B200_LIB=$PWD/variants/nostream.so python bench.py --quick --no-cpu --scene mega --steps 3 --warmup 1 2>>gpurun_out/ab_mega.err | tail -1 | python -c "$summ" nostream
echo "== ncu instruction counts (one step)"
python bench.py --one-step --scene cornell_monkey > gpurun_out/onestep_plain.log 2>&1 && \
ncu --profile-from-start off --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none --csv --log-file gpurun_out/inst_cornell_monkey.csv python bench.py --one-step --scene cornell_monkey > gpurun_out/onestep_ncu.log 2>&1
echo "ncu exit $?"; wc -l gpurun_out/inst_cornell_monkey.csv
