#!/bin/bash
# resident tree kernel CTA size with the S16 stack (768 threads / 80 registers against 896 / 72 and 1024 / 64)
summ='import json,sys
d=json.loads(sys.stdin.read())
print("%-14s %-12s Mrays/s %7.1f  ms/step %7.3f  stages %s" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}))'
for sc in cornell_monkey cornell_boxes matball; do
python bench.py --quick --no-cpu --scene $sc 2>/dev/null | tail -1 | python -c "$summ" $sc default
for so in variants/*.so; do
  PTINA_B200_LIB=$PWD/$so python bench.py --quick --no-cpu --scene $sc 2>/dev/null | tail -1 | python -c "$summ" $sc $(basename $so .so)
done
done
