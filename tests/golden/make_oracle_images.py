"""Converged reference images at BASELINE.json's stated sizes, rendered by the CPU oracle (oracle/ptina_oracle.cpp, which
tests/test_golden.py pins bit for bit to the reference's own sources executed under the shim).  The GPU suite renders the same Sobol
indices and asserts north_star's image gate -- per-pixel relative RMSE <= 1e-3 -- at full size (tests/test_gpu_fullsize.py).

    python tests/golden/make_oracle_images.py [--only NAME] [--threads N]

  image_cornell_boxes.npz   config 1: 34 triangles, 512x512, PathEngine, Sobol points 65..1088 (1024 spp)      ~4 min on 8 cores
  image_cornell_monkey.npz  config 2: 978 triangles, 512x512, PathEngine, Sobol points 65..1088 (1024 spp)     ~5 min
  image_matball.npz         config 3: 1024x1024, BruteEngine, textures + environment, Sobol points 65..128 (64 spp); every second
                            pixel in x and y is stored (262 144 pixels: the gate is a mean over pixels)
Stored: float32 `rgb` = sum / w exactly as FilmTable.get_image (filmtable.py:47-63), plus the Sobol range and the oracle's ray count.
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from ptina_b200 import scenes  # noqa: E402

JOBS = {
    'cornell_boxes': dict(spp=1024, stride=1),
    'cornell_monkey': dict(spp=1024, stride=1),
    'matball': dict(spp=64, stride=2),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='')
    ap.add_argument('--threads', type=int, default=0)
    a = ap.parse_args()
    for name, job in JOBS.items():
        if a.only and a.only != name:
            continue
        sc = scenes.CONFIGS[name]()
        o = oracle.Oracle()
        scenes.apply(o, sc)
        eng = oracle.ENGINE_BRUTE if sc['engine'] == 'brute' else oracle.ENGINE_PATH
        t0 = time.time()
        k_first = o.sobol_time + 1
        cnt = o.render(eng, job['spp'], nthreads=a.threads)
        img = o.get_image()
        st = job['stride']
        np.savez_compressed(os.path.join(HERE, f'image_{name}.npz'), rgb=img[::st, ::st, :3].copy(), k_first=np.int32(k_first), spp=np.int32(job['spp']),
                            stride=np.int32(st), size=np.asarray(sc['size'], np.int32), rays=np.int64(cnt['rays']))
        print(f'{name}: {job["spp"]} spp from k={k_first} in {time.time() - t0:.1f}s, {cnt["rays"]} rays', flush=True)


if __name__ == '__main__':
    main()
