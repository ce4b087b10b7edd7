"""Config 2 (and others): does splitting ONE batch into concurrent chunks on several lanes pay?  (run under gpurun)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptina_b200 import _native, scenes, worker
scene = sys.argv[1] if len(sys.argv) > 1 else 'cornell_monkey'
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sc = scenes.CONFIGS[scene]()
worker.init(); ctx = _native.context(); scenes.apply(worker, sc)
eng = _native.ENGINE_BRUTE if sc['engine'] == 'brute' else _native.ENGINE_PATH
def run(label):
    ctx.sobol_reset(); worker.clear()
    for _ in range(3): ctx.render_range(eng, 65, spp, 1)
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ctx.render_range(eng, 65, spp, 1)
    ctx.flush(); e1.record(); torch.cuda.synchronize()
    print(json.dumps({'scene': scene, 'setting': label, 'ms_per_step': round(e0.elapsed_time(e1) / 10, 3)}), flush=True)
run('one batch, one lane')
ctx.set_option('pt_split', 1)
for k in (2, 3, 4):
    ctx.set_option('pt_lanes', k); run(f'one batch split over {k} lanes')
ctx.set_option('pt_split', 0); ctx.set_option('pt_lanes', 2)
