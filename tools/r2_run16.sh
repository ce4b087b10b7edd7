#!/bin/bash
summ='import json,sys
d=json.loads(sys.stdin.read())
print("%-14s Mrays/s %7.1f  ms/step %7.3f  e2e %7.1f stages %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}))'
python -m pytest tests -q -m gpu -x > gpurun_out/tests_default.log 2>&1; tail -15 gpurun_out/tests_default.log
for sc in cornell_monkey cornell_boxes matball; do python bench.py --quick --no-cpu --scene $sc 2>/dev/null | tail -1 | python -c "$summ" $sc; done
python bench.py --quick --no-cpu --scene mega --steps 3 --warmup 1 2>/dev/null | tail -1 | python -c "$summ" mega
