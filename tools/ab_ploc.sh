run() {
python bench.py --scene $1 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
c=d['counters_per_step_rank0']
print('$2', d['config']['scene'], 'Mrays/s %.1f  ms/step %.3f  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), 'nodes/ray %.2f tris/ray %.2f' % (c['node_visits']/c['rays'], c['tri_tests']/c['rays']))
"
}
for v in b4 b2 b64; do export PTINA_B200_LIB=$PWD/variants/$v.so; for sc in cornell_monkey cornell_boxes matball; do run $sc $v; done; done
