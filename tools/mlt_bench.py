"""Config 5 (BASELINE.json configs[4], exams/metropolis.py): MLTPathEngine defaults -- 2^18 chains x 32 dimensions per render(),
LSP 0.25, sigma 0.01 -- on the config-2 scene at 512x512.  Prints device-timed chain proposals/s and Mrays/s.  Run under gpurun."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptina_b200 import _native, scenes, worker
from ptina_b200.engine import MLTPathEngine

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 64
lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sc = scenes.CONFIGS['metropolis']()
worker.init()
ctx = _native.context()
ctx.use_torch_stream()
scenes.apply(worker, sc)
if lanes:
    ctx.set_option('mlt_lanes', lanes)
nch = 1 << 18
eng = MLTPathEngine(nchains=nch, seed=0)
eng.LSP[None] = 0.25; eng.Sigma[None] = 0.01
eng.reset(); worker.clear()
eng.render(8)                                   # warm-up (also past the all-accept first iteration)
ctx.set_counting(True, False); ctx.reset_counters()
eng.render(4); ctx.synchronize()
rays_per_iter = ctx.counters()['rays'] / 4
ctx.set_counting(False, False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
eng.render(iters)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(json.dumps({'lanes': lanes or 'default (2)', 'workload': f'metropolis: 978 tris, 512x512, {nch} chains x 32 dims per render()', 'ms_per_render': ms,
                  'proposals_per_s': nch / (ms * 1e-3), 'rays_per_render': rays_per_iter, 'Mrays_per_s': rays_per_iter / (ms * 1e-3) / 1e6}))
