"""Diagnostics (GPU box): per-policy counters, stage times and film equality on one scene."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ptina_b200 import scenes, worker, _native

name = sys.argv[1] if len(sys.argv) > 1 else 'cornell_monkey'
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
worker.init()
ctx = _native.context()
sc = scenes.CONFIGS[name]()
scenes.apply(worker, sc)
films = {}
for pol, label in ((_native.TRAVERSE_ORDERED, 'ordered'), (_native.TRAVERSE_ORDERED_EXACT, 'exact'), (_native.TRAVERSE_REFERENCE, 'reference')):
    ctx.set_traversal(pol)
    ctx.sobol_reset(); worker.clear()
    ctx.set_counting(True, True); ctx.reset_counters()
    ctx.render(_native.ENGINE_PATH if sc['engine'] == 'path' else _native.ENGINE_BRUTE, spp); ctx.synchronize()
    print(label, ctx.counters(), {k: round(v, 3) for k, v in ctx.stage_ms().items()})
    ctx.set_counting(False, False)
    films[label] = ctx.get_film().copy()
for k in ('exact', 'reference'):
    print('film ordered ==', k, np.array_equal(films['ordered'], films[k]), float(np.abs(films['ordered'] - films[k]).max()))
