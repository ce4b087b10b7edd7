run() {
python bench.py --scene $1 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
c=d['counters_per_step_rank0']
print('$2', d['config']['scene'], 'Mrays/s %.1f  ms/step %.3f  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), 'nodes/ray %.2f build %.3f' % (c['node_visits']/c['rays'], d['tree']['build_ms']))
"
}
for v in r8 r32; do export PTINA_B200_LIB=$PWD/variants/$v.so; for sc in cornell_monkey matball; do run $sc $v; done; done
