"""north_star's image gate at full length: converged 1024-spp image of a config (default config 2, 512x512) from the CUDA
path vs the CPU oracle with the same Sobol indices -- per-pixel relative RMSE <= 1e-3.  Takes a few minutes of CPU time on
the GPU box (the oracle is the slow side); not part of the test suite for that reason."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from ptina_b200 import scenes, worker, _native

name = sys.argv[1] if len(sys.argv) > 1 else 'cornell_monkey'
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
sc = scenes.CONFIGS[name]()
worker.init(); ctx = _native.context()
scenes.apply(worker, sc)
eng = _native.ENGINE_BRUTE if sc['engine'] == 'brute' else _native.ENGINE_PATH
t = time.time(); ctx.render(eng, spp); ctx.synchronize(); tg = time.time() - t
a = worker.get_image()[..., :3]
ref = oracle.Oracle(); scenes.apply(ref, sc)
t = time.time(); ref.render(oracle.ENGINE_BRUTE if sc['engine'] == 'brute' else oracle.ENGINE_PATH, spp); tc = time.time() - t
b = ref.get_image()[..., :3]
rel = (a - b) / np.maximum(b, 1e-2)
rmse = float(np.sqrt((rel ** 2).mean()))
print(f'{name} {sc["size"]} {spp} spp: rel-RMSE {rmse:.3e}  max |rel| {np.abs(rel).max():.3e}  pixels with |rel| > 1e-3: {(np.abs(rel).max(2) > 1e-3).sum()}  '
      f'gpu {tg:.2f} s  oracle {tc:.1f} s ({oracle.num_threads()} threads)')
assert rmse <= 1e-3
