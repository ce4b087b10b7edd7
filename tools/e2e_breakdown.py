import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ptina_b200 import scenes, worker, _native
from ptina_b200.model import ModelPool
from ptina_b200.tree import BVHTree
sc = scenes.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'cornell_monkey']()
if sc['name'] == 'mega': sc['spp'] = 4
worker.init(); ctx = _native.context(); scenes.apply(worker, sc)
nx, ny = sc['size']
vp = torch.from_numpy(np.ascontiguousarray(sc['vertices'], dtype=np.float32)).pin_memory()
mp = torch.from_numpy(np.ascontiguousarray(sc['mtlids'], dtype=np.int32)).pin_memory()
img = torch.empty((nx, ny, 4), dtype=torch.float32).pin_memory()
def T(name, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
    acc.setdefault(name, []).append(dt)
acc = {}
for it in range(6):
    T('load', lambda: ModelPool().load(vp.numpy(), mp.numpy()))
    T('build', lambda: BVHTree().build())
    T('clear', lambda: worker.clear())
    T('render', lambda: ctx.render(_native.ENGINE_PATH, sc['spp']))
    T('get_image', lambda: ctx.get_image(0, out=img.numpy()))
    T('film_tensor', lambda: ctx.film_tensor(0))
for k, v in acc.items():
    print(f'{k:12s} ' + ' '.join(f'{x:8.3f}' for x in v))
print('tree build_ms (device):', ctx.tree.build_ms, 'sweeps', ctx.tree.aabb_sweeps)
