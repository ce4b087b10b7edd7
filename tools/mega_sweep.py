"""Config 4 (1.02 M triangles, 1920x1080): traversal-tree builder and lane schedule sweep in one process (run under gpurun).
Prints node visits / triangle tests per ray, build time, and device-timed Mrays/s of a 32-spp step for each setting."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ptina_b200 import _native, scenes, worker
from ptina_b200.tree import BVHTree

scene = sys.argv[1] if len(sys.argv) > 1 else 'mega'
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sc = scenes.CONFIGS[scene]()
worker.init()
ctx = _native.context()
scenes.apply(worker, sc)
eng = _native.ENGINE_BRUTE if sc['engine'] == 'brute' else _native.ENGINE_PATH


def measure(label):
    BVHTree().build()
    info = ctx.tree
    ctx.sobol_reset(); worker.clear()
    ctx.set_counting(True, False); ctx.reset_counters()
    ctx.render_range(eng, 65, 4, 1); ctx.synchronize()
    c = ctx.counters()
    ctx.set_counting(False, False)
    ctx.render_range(eng, 65, 8, 1); ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ctx.render_range(eng, 65, spp, 1); ctx.flush(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rays = c['rays'] / 4 * spp
    print(json.dumps({'setting': label, 'Mrays_per_s': round(rays / ms / 1e3, 1), 'ms': round(ms, 2), 'nodes_per_ray': round(c['node_visits'] / c['rays'], 2),
                      'tris_per_ray': round(c['tri_tests'] / c['rays'], 2), 'build_ms': round(info.build_ms, 2), 'trav_depth': info.trav_depth, 'ploc': info.trav_ploc}), flush=True)


for wide in (0, 1):
    ctx.set_option('wide4', wide)
    measure(f'ploc r4, 2 lanes, wide4={wide}')
ctx.set_option('wide4', 1)
ctx.set_option('ploc_big', 0); measure('lbvh topology, wide4=1'); ctx.set_option('ploc_big', 1)
ctx.set_option('pt_lanes', 1); measure('ploc r4, 1 lane, wide4=1'); ctx.set_option('pt_lanes', 2)
