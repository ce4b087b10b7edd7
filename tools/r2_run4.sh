#!/bin/bash
python -m pytest tests -q -m gpu > gpurun_out/tests_default.log 2>&1; tail -4 gpurun_out/tests_default.log; grep -h "rel-RMSE\|fast mode" gpurun_out/tests_default.log
python -m pytest tests/test_gpu_fullsize.py -q -m gpu -s 2>&1 | grep -h "rel-RMSE\|fast mode\|matball:"
echo "== MLT one lane vs two lanes"
PTB_MLT_ONE_LANE=1 python tools/mlt_bench.py 2>&1 | tail -1
python tools/mlt_bench.py 2>&1 | tail -1
echo "== bench"
python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "exit $?"; tail -c 600 gpurun_out/bench_r2b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2b.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'percall', d['e2e_percall']['value'], 'sustained', d['sustained']['value'])
print('roofline', {k:v for k,v in d['roofline'].items() if k in ('bound','achieved','peak','frac','lanes_per_inst','frac_lane_weighted','traffic')})
for k,v in d['configs'].items():
    print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('Mrays_per_s','spp_per_s','ms_per_step','Mproposals_per_s','ms_per_render')}, v.get('roofline_l2',{}).get('frac'), v.get('stage_ms_per_step'))
PY
