"""Launch list of BVHTree().build() on the config-2 scene (run under ncu --metrics gpu__time_duration.sum): where the 0.29 ms of the e2e step go."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ptina_b200 import scenes, worker, _native
from ptina_b200.tree import BVHTree
name = sys.argv[1] if len(sys.argv) > 1 else 'cornell_monkey'
worker.init()
ctx = _native.context()
sc = scenes.CONFIGS[name]()
scenes.apply(worker, sc)
ctx.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter(); BVHTree().build(); ts.append((time.perf_counter() - t0) * 1e3)
print('host ms per build', [round(t, 3) for t in ts], 'device ms', round(ctx.tree.build_ms, 3))
