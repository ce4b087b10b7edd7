#!/bin/bash
# 2-GPU call: multi-GPU correctness (two frames, out-of-place reduce), MLT on two lanes, then the full bench at N=2
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "mlt or batching or overlapped" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/dist_check.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -6
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo "exit $?"; tail -c 1500 gpurun_out/bench_r2_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_n2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'film_check', d.get('film_check'), 'render/reduce ms', d.get('render_ms'), d.get('reduce_ms'))
for k,v in d['configs'].items():
    print(k, {kk: (round(vv,3) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('Mrays_per_s','spp_per_s','ms_per_step','Mproposals_per_s','ms_per_render','render_ms','reduce_ms','host_enqueue_ms_per_step')}, v.get('film_check'), v.get('limits'), v.get('full_1024spp',{}).get('seconds'))
PY
PTB_MLT_ONE_LANE=1 python tools/mlt_bench.py 2>&1 | tail -2
python tools/mlt_bench.py 2>&1 | tail -2
