"""The `ptina` / `taichi` compatibility packages and the generated assets: what an unmodified PTina driver
script needs before it reaches the GPU.  The GPU half (actually running exams/benchmark.py) is tests/test_gpu_fullsize.py."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMS = os.path.join(ROOT, 'tests', 'golden', 'exams')


@pytest.mark.parametrize('name', ['benchmark.py', 'benchtiles.py'])
def test_driver_script_fixture_is_the_reference_text(name):
    ref = os.path.join('/root/reference/exams', name)
    if not os.path.exists(ref):
        pytest.skip('reference checkout not present (GPU box)')
    sha = lambda p: hashlib.sha256(open(p, 'rb').read()).hexdigest()
    assert sha(ref) == sha(os.path.join(EXAMS, name + '.txt'))


def test_compat_names_without_gpu():
    """`from ptina.things import *` etc. give the names the reference's modules give (in a subprocess: the packages are not
    importable unless the compat directory is put on the path)."""
    code = '''
import sys
from ptina_b200 import compat
assert 'ptina' not in sys.modules and 'taichi' not in sys.modules or 'compat' in sys.modules['taichi'].__file__
compat.install()
ns = {}
exec("from ptina.things import *\\nfrom ptina.engine.path import *\\nfrom ptina.tools.readgltf import readgltf", ns)
for k in ('ti', 'np', 'init_things', 'PathEngine', 'FilmTable', 'ModelPool', 'MaterialPool', 'ImagePool', 'BVHTree', 'Camera', 'LightPool', 'WorldLight', 'SobolSampler', 'readgltf', 'wanghash2'):
    assert k in ns, k
exec("from ptina.engine.brute import *\\nfrom ptina.engine.mltpath import *\\nfrom ptina.engine.preview import *\\nfrom ptina.tools.readobj import readobj\\nfrom ptina.tools.mtworker import *\\nimport ptina.worker", ns)
for k in ('BruteEngine', 'MLTPathEngine', 'PreviewEngine', 'readobj', 'DaemonModule', 'OnDemandProxy'):
    assert k in ns, k
ti = ns['ti']
ti.init(ti.cuda)
assert ti.init_args['arch'] == 'cuda' and ti.opengl == 'opengl'
import numpy as np
img = np.arange(2 * 3 * 3, dtype=np.float32).reshape(2, 3, 3)
ti.imshow(img, 'x')
assert ti.shown == [('x', (2, 3, 3))] and ti.imresize(img, 4).shape == (4, 4, 3) and ti.imresize(img, 4, 6)[3, 5, 0] == img[1, 2, 0]
assert callable(ns['PathEngine'].render_tile) and callable(ns['PathEngine'].render_final)
print('ok')
'''
    r = subprocess.run([sys.executable, '-c', code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip() == 'ok', r.stderr


def test_assets_are_regenerable_and_match_the_configs(tmp_path):
    from ptina_b200 import scenes
    from ptina_b200.tools.make_assets import make_assets
    from ptina_b200.tools.readgltf import readgltf
    from ptina_b200.tools.readobj import readobj
    names = make_assets(str(tmp_path))
    assert names == ['cornell.gltf', 'monkey_cornell.gltf', 'uvsphere.obj']
    for n in names:
        assert open(tmp_path / n, 'rb').read() == open(os.path.join(ROOT, 'assets', n), 'rb').read(), f'assets/{n} is stale: python -m ptina_b200.tools.make_assets'
    v, m, mats, imgs = readgltf(os.path.join(ROOT, 'assets', 'monkey_cornell.gltf'))
    sc = scenes.cornell_monkey()
    assert v.shape == (978 * 3, 8) and imgs == [] and len(mats) == 4 and np.bincount(m).tolist() == [6, 2, 2, 968]
    order = np.argsort(sc['mtlids'], kind='stable')
    want = np.asarray(sc['vertices'], np.float32).reshape(-1, 24)[order].reshape(-1, 8)
    assert np.array_equal(v[:, :3].astype(np.float32), want[:, :3]) and np.abs(v - want).max() < 1e-6
    assert np.allclose(mats[3][0][0][:3], [0.85, 0.6, 0.25]) and mats[3][1] == (pytest.approx(0.35), -1) and len(mats[0]) == 3
    o = readobj(os.path.join(ROOT, 'assets', 'uvsphere.obj'))
    assert o['f'].shape == (960, 3, 3) and np.allclose(np.linalg.norm(o['v'], axis=1), 1, atol=1e-6)
