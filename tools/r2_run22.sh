#!/bin/bash
# resident tree kernel: node steps per vote, pending-leaf ring size, CTA size; quantised-resident vs Node64 (config 2, 1, 3)
summ='import json,sys
d=json.loads(sys.stdin.read())
print("%-14s %-12s Mrays/s %7.1f  ms/step %7.3f  stages %s" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}))'
for sc in cornell_monkey matball; do
for rep in 1 2; do python bench.py --quick --no-cpu --scene $sc 2>/dev/null | tail -1 | python -c "$summ" $sc default; done
for rep in 1 2; do PTB_QUANT_RESIDENT_BVH=1 python bench.py --quick --no-cpu --scene $sc 2>/dev/null | tail -1 | python -c "$summ" $sc quant-res; done
for so in variants/*.so; do
  PTINA_B200_LIB=$PWD/$so python bench.py --quick --no-cpu --scene $sc 2>/dev/null | tail -1 | python -c "$summ" $sc $(basename $so .so)
done
done
for rep in 1 2; do python bench.py --quick --no-cpu --scene cornell_boxes 2>/dev/null | tail -1 | python -c "$summ" cornell_boxes default; PTB_QUANT_RESIDENT_BVH=1 python bench.py --quick --no-cpu --scene cornell_boxes 2>/dev/null | tail -1 | python -c "$summ" cornell_boxes quant-res; done
