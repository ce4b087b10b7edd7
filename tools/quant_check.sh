run() {
python bench.py --scene $1 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$2', d['config']['scene'], 'Mrays/s %.1f  ms/step %.3f  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), {k: round(v,3) for k,v in d['stage_ms_per_step'].items()})
"
}
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
run cornell_monkey new; run cornell_boxes new; run matball new; run mega new
