#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; echo "exit $?"; tail -c 800 gpurun_out/bench_r2_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n8.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'film_check', d.get('film_check'), 'render/reduce ms', d.get('render_ms'), d.get('reduce_ms'))
for k,v in d['configs'].items():
    print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('Mrays_per_s','spp_per_s','ms_per_step','Mproposals_per_s','ms_per_render','render_ms','reduce_ms','host_enqueue_ms_per_step')}, (v.get('film_check') or {}).get('max_rel_diff'), v.get('limits'), v.get('full_1024spp',{}).get('seconds'))
PY
