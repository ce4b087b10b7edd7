#!/bin/bash
# bench every tuning build under variants/ (and the default library) -- run under gpurun
summ='import json,sys
d=json.loads(sys.stdin.read())
print("%-10s Mrays/s %7.1f  ms/step %6.2f  stages %s" % (sys.argv[1], d["value"], d["ms_per_step"], {k: round(v,2) for k,v in d["stage_ms_per_step"].items()}))'
python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu ${BENCH_ARGS} 2>/dev/null | tail -1 | python -c "$summ" default
PTB_NO_RESIDENT_BVH=1 python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu ${BENCH_ARGS} 2>/dev/null | tail -1 | python -c "$summ" default-global
for so in variants/*.so; do
  PTINA_B200_LIB=$PWD/$so python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu ${BENCH_ARGS} 2>/dev/null | tail -1 | python -c "$summ" $(basename $so .so)
  if [[ $so == *m6* ]]; then PTB_NO_RESIDENT_BVH=1 PTINA_B200_LIB=$PWD/$so python bench.py --steps ${STEPS:-5} --warmup 3 --no-cpu ${BENCH_ARGS} 2>/dev/null | tail -1 | python -c "$summ" $(basename $so .so)-global; fi
done
