run() {
python bench.py --scene $1 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$2', d['config']['scene'], 'Mrays/s %.1f  ms/step %.3f  e2e %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value']), 'stages sum %.3f' % sum(d['stage_ms_per_step'].values()))
"
}
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
run cornell_monkey overlap; PTB_NO_OVERLAP=1 run cornell_monkey serial
run cornell_boxes overlap; PTB_NO_OVERLAP=1 run cornell_boxes serial
run mega overlap; PTB_NO_OVERLAP=1 run mega serial
