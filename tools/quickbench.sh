#!/bin/bash
# tests + both benches, compact output (run under gpurun)
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for sc in cornell_monkey mega; do
python bench.py --scene $sc --steps ${STEPS:-5} --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['config']['scene'], 'Mrays/s %.1f  ms/step %.2f  spp/s %.1f  e2e %.1f' % (d['value'], d['ms_per_step'], d['spp_per_s'], d['e2e']['value']))
print('  stage', {k: round(v,3) for k,v in d['stage_ms_per_step'].items()}, 'roofline GB/s %.0f' % d['roofline']['achieved'])
c=d['counters_per_step_rank0']; print('  per ray: nodes %.1f boxes %.1f tris %.1f' % (c['node_visits']/c['rays'], c['box_tests']/c['rays'], c['tri_tests']/c['rays']))
"
done
