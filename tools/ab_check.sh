#!/bin/bash
# quick A/B helper (run under gpurun): GPU tests in the three node modes, then the four bench scenes, compact output.
# PTINA_B200_LIB=/path/to/variant.so selects a tuning build (python -m ptina_b200.build -DNAME=v --out=...).
run() {
python bench.py --scene $1 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['scene'], 'Mrays/s %.1f  ms/step %.3f  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(v,3) for k,v in d['stage_ms_per_step'].items()})
"
}
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
PTB_QUANT_RESIDENT_BVH=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
PTB_NO_RESIDENT_BVH=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for sc in cornell_monkey cornell_boxes matball mega; do run $sc; done
