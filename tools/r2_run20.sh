#!/bin/bash
# tuning sweep of the global-memory tree kernel on config 4 (4-wide nodes, PLOC tree, two lanes)
summ='import json,sys
d=json.loads(sys.stdin.read())
print("%-10s %-10s Mrays/s %7.1f  ms/step %7.3f  stages %s" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], {k: round(v,2) for k,v in d["stage_ms_per_step"].items()}))'
python bench.py --quick --no-cpu --scene mega --steps 3 --warmup 1 2>/dev/null | tail -1 | python -c "$summ" mega default
for so in variants/*.so; do
  PTINA_B200_LIB=$PWD/$so python bench.py --quick --no-cpu --scene mega --steps 3 --warmup 1 2>/dev/null | tail -1 | python -c "$summ" mega $(basename $so .so)
done
python bench.py --quick --no-cpu --scene mega --steps 3 --warmup 1 2>/dev/null | tail -1 | python -c "$summ" mega default-again
