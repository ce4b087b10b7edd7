run() {
python bench.py --scene $1 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
c=d['counters_per_step_rank0']
print('$2', d['config']['scene'], 'Mrays/s %.1f  ms/step %.3f  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(v,3) for k,v in d['stage_ms_per_step'].items()}, 'nodes/ray %.2f tris/ray %.2f build %.3f' % (c['node_visits']/c['rays'], c['tri_tests']/c['rays'], d['tree']['build_ms']))
"
}
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for sc in cornell_monkey cornell_boxes matball; do run $sc frag; PTB_NO_FRAGMENTS=1 run $sc list; done
