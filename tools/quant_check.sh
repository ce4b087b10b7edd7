run() {
python bench.py --scene $1 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$2', d['config']['scene'], 'Mrays/s %.1f  ms/step %.2f' % (d['value'], d['ms_per_step']))
print('  stage', {k: round(v,3) for k,v in d['stage_ms_per_step'].items()})
c=d['counters_per_step_rank0']; print('  per ray: nodes %.2f boxes %.2f tris %.2f' % (c['node_visits']/c['rays'], c['box_tests']/c['rays'], c['tri_tests']/c['rays']))
"
}
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
PTB_QUANT_RESIDENT_BVH=1 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
run matball default; run cornell_monkey default
PTB_QUANT_RESIDENT_BVH=1 run cornell_monkey quant-resident
PTB_NO_RESIDENT_BVH=1 run matball global
