import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np
import oracle
from ptina_b200 import worker, _native
from ptina_b200.model import ModelPool
from ptina_b200.tree import BVHTree
import test_gpu_parity as T
worker.init(); gpu = _native.context()
kind = sys.argv[1] if len(sys.argv) > 1 else 'few_big'
for seed in range(40):
    rng = np.random.default_rng(1000 * seed + {'few_big': 5, 'many_big': 6, 'no_big': 7}[kind])
    verts = T._soup(rng, kind); nf = verts.shape[0] // 3
    ModelPool().load(verts, np.zeros(nf, np.int32))
    try: BVHTree().build()
    except RuntimeError: continue
    if gpu.tree.valid: break
o = oracle.Oracle(); o.load_model(verts, np.zeros(nf, np.int32)); o.build_tree()
print('seed', seed, 'n', nf, 'list', gpu.tree.list_n, 'policy', gpu.tree.policy)
tri = verts[:, :3].reshape(nf, 3, 3)
m = 30000
f = rng.integers(0, nf, m); w = rng.dirichlet([1, 1, 1], m).astype(np.float32)
tgt = (tri[f] * w[:, :, None]).sum(1); sel = rng.integers(0, 4, m); k = rng.integers(0, 3, m)
tgt = np.where((sel == 1)[:, None], tri[f, k], tgt)
mid = (tri[f, k] + tri[f, (k + 1) % 3]) * np.float32(0.5)
tgt = np.where((sel == 2)[:, None], mid, tgt)
out = mid + (mid - tri[f, (k + 2) % 3]) * (10.0 ** rng.uniform(-7, -3, (m, 1))).astype(np.float32)
tgt = np.where((sel == 3)[:, None], out, tgt).astype(np.float32)
org = rng.uniform(-3, 3, (m, 3)).astype(np.float32)
org[: m // 4] = (tri[rng.integers(0, nf, m // 4)] * rng.dirichlet([1, 1, 1], m // 4).astype(np.float32)[:, :, None]).sum(1)
d = tgt - org; d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
d[m // 2: m // 2 + 2000, rng.integers(0, 3)] *= np.float32(1e-4)
rays = np.ascontiguousarray(np.concatenate([org, d], 1), np.float32)
avoid = np.where(rng.random(m) < 0.3, f, -1).astype(np.int32)
ref = o.intersect(rays, avoid); h = ref['hit'] == 1
dis = np.where(h, ref['depth'] * rng.choice([0.5, 1.0, 1.0, 1.5], m), 5.0).astype(np.float32)
want = (h & (ref['depth'] <= dis)).astype(np.int32)
for pol in (0, 3, 1):
    got = gpu.occluded(rays, dis, avoid, pol)
    bad = np.nonzero(got != want)[0]
    print('policy', pol, 'bad', bad.size)
    for i in bad[:12]:
        print('  ray', i, rays[i], 'avoid', avoid[i], 'dis', dis[i], 'ref depth', ref['depth'][i], 'ref idx', ref['index'][i], 'got', got[i], 'want', want[i])
