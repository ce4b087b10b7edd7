"""Probe: how much would two independent half-batches per frame (two wavefronts on two streams) gain?
Two native contexts on one GPU stand in for the two lanes.  Run under gpurun."""
import sys
import torch
from ptina_b200 import _native, scenes, worker, things

name = sys.argv[1] if len(sys.argv) > 1 else 'cornell_monkey'
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32
sc = scenes.CONFIGS[name]()
ctxs = []
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for i in range(2):
    worker.init()
    c = _native.context()
    scenes.apply(worker, sc)
    with torch.cuda.stream(streams[i]):
        c.use_torch_stream()
    ctxs.append(c)
    _native._ctx = None
    from ptina_b200.sampling.sobol import SobolSampler
    from ptina_b200 import engine
    for cls in things._POOLS + (SobolSampler,) + engine.ENGINES:
        cls._forget()
A, B = ctxs
eng = _native.ENGINE_PATH


def run(label, fn, steps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.default_stream())
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps * 1e3
    print(f'{label:28s} {dt:8.3f} ms/step')


k = [65]
def single():
    A.render_range(eng, k[0], spp, 1); k[0] += spp
def lanes():
    A.render_range(eng, k[0], spp // 2, 1); B.render_range(eng, k[0] + spp // 2, spp // 2, 1); k[0] += spp
def halves_serial():
    A.render_range(eng, k[0], spp // 2, 1); A.render_range(eng, k[0] + spp // 2, spp // 2, 1); k[0] += spp
run('one wavefront', single)
run('two half wavefronts, serial', halves_serial)
run('two half wavefronts, 2 ctx', lanes)
run('one wavefront', single)
