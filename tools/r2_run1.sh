#!/bin/bash
# round-2 GPU call 1: whole GPU suite, A/B of the list filter and the fast-shade build, then the full bench (all configs)
summ='import json,sys
d=json.loads(sys.stdin.read())
print("%-12s Mrays/s %7.1f  ms/step %6.3f  e2e %7.1f  percall %s  stages %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], (d.get("e2e_percall") or {}).get("value"), {k: round(v,3) for k,v in d["stage_ms_per_step"].items()}))'
python -m pytest tests -q -m gpu -x 2>&1 | tail -15
echo "== A/B (config 2, --quick)"
python bench.py --quick --no-cpu 2>gpurun_out/ab_default.err | tail -1 | python -c "$summ" default
for v in nofilter fastshade; do
  PTINA_B200_LIB=$PWD/variants/$v.so python bench.py --quick --no-cpu 2>gpurun_out/ab_$v.err | tail -1 | python -c "$summ" $v
done
echo "== A/B mega (--quick)"
python bench.py --quick --no-cpu --scene mega --steps 3 --warmup 1 2>>gpurun_out/ab_default.err | tail -1 | python -c "$summ" default
PTINA_B200_LIB=$PWD/variants/nofilter.so python bench.py --quick --no-cpu --scene mega --steps 3 --warmup 1 2>>gpurun_out/ab_nofilter.err | tail -1 | python -c "$summ" nofilter
echo "== GPU suite on the fast-shade build (which gates hold?)"
PTINA_B200_LIB=$PWD/variants/fastshade.so python -m pytest tests -q -m gpu 2>&1 | tail -25
echo "== full bench"
python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "exit $?"; tail -c 600 gpurun_out/bench_r2a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2a.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'percall', d['e2e_percall']['value'], 'sustained', d['sustained']['value'], d['sustained']['clocks'])
for k,v in d['configs'].items():
    print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('Mrays_per_s','spp_per_s','ms_per_step','Mproposals_per_s','ms_per_render')}, v.get('roofline_l2',{}).get('frac'), v.get('full_1024spp'))
print('cpu', d.get('cpu_baseline'))
PY
