#!/bin/bash
# parity suites in the three node modes + the quick benches
summ='import json,sys
d=json.loads(sys.stdin.read())
print("%-14s Mrays/s %7.1f  ms/step %7.3f  e2e %7.1f stages %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}))'
python -m pytest tests -q -m gpu -x > gpurun_out/tests_default.log 2>&1; tail -3 gpurun_out/tests_default.log
PTB_QUANT_RESIDENT_BVH=1 python -m pytest tests -q -m gpu -x > gpurun_out/tests_qres.log 2>&1; tail -3 gpurun_out/tests_qres.log
PTB_NO_RESIDENT_BVH=1 python -m pytest tests -q -m gpu -x > gpurun_out/tests_global.log 2>&1; tail -3 gpurun_out/tests_global.log
for sc in cornell_monkey matball; do python bench.py --quick --no-cpu --scene $sc 2>/dev/null | tail -1 | python -c "$summ" $sc; done
PTB_QUANT_RESIDENT_BVH=1 python bench.py --quick --no-cpu --scene cornell_monkey 2>/dev/null | tail -1 | python -c "$summ" c2-quant-resident
python bench.py --quick --no-cpu --scene mega --steps 3 --warmup 1 2>/dev/null | tail -1 | python -c "$summ" mega
