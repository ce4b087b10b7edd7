#!/bin/bash
# shared-memory traversal stack: correctness in all three node modes, then the depth sweep
summ='import json,sys
d=json.loads(sys.stdin.read())
print("%-14s %-10s Mrays/s %7.1f  ms/step %7.3f  stages %s" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}))'
python -m pytest tests -q -m gpu > gpurun_out/tests_default.log 2>&1; tail -3 gpurun_out/tests_default.log
PTB_NO_RESIDENT_BVH=1 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu 2>&1 | tail -2
PTB_QUANT_RESIDENT_BVH=1 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu 2>&1 | tail -2
run() { PTINA_B200_LIB=$2 python bench.py --quick --no-cpu --scene $1 --steps $3 --warmup $4 2>/dev/null | tail -1 | python -c "$summ" $1 $5; }
for sc in cornell_monkey matball; do
  run $sc $PWD/ptina_b200/libptina_b200.so 10 3 s8g8
  for v in s0g0 s4g4 s6g6 s12g12; do run $sc $PWD/variants/$v.so 10 3 $v; done
done
run mega $PWD/ptina_b200/libptina_b200.so 3 1 s8g8
for v in s0g0 s4g4 s6g6 s12g12 s8g16; do run mega $PWD/variants/$v.so 3 1 $v; done
