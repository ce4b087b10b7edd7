"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol the header
declares; the facade exposes the reference's names; nothing in the product imports the oracle; no CPU fallback."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'ptina_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(ptb_[a-z0-9_]+)\s*\(', text)))


def test_header_symbols_exported(native_so):
    from ptina_b200 import _native
    names = header_symbols()
    assert len(names) >= 40
    lib = ctypes.CDLL(native_so)
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/ptina_b200.h but not exported by libptina_b200.so'
    assert sorted(_native.SYMBOLS) == names, 'ctypes binding and header disagree'


def test_library_is_sm100a(native_so):
    out = subprocess.run(['cuobjdump', '--list-elf', native_so], capture_output=True, text=True).stdout
    assert 'sm_100a' in out and 'sm_90' not in out


def test_no_cpu_fallback(native_so):
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from ptina_b200 import _native
    with pytest.raises(_native.NativeError, match='no CUDA device'):
        _native.Context(device=0)


def test_facade_mirrors_reference_names():
    # modules import without touching the GPU; names follow ptina/worker.py:11-87 and the singleton classes
    from ptina_b200 import worker, things, multimesh
    for fn in ('init', 'synchronize', 'render', 'render_preview', 'set_size', 'get_size', 'clear', 'set_mlt_param', 'get_image',
               'fast_export_image', 'clear_lights', 'set_world_light', 'add_light', 'load_model', 'load_images', 'load_materials',
               'build_tree', 'set_camera'):
        assert callable(getattr(worker, fn)), fn
    assert callable(things.init_things) and callable(multimesh.compose_multiple_meshes)
    from ptina_b200.engine import PathEngine, BruteEngine, PreviewEngine, MLTPathEngine   # noqa: F401
    from ptina_b200.tree import BVHTree                                                    # noqa: F401
    from ptina_b200.model import ModelPool                                                 # noqa: F401
    from ptina_b200.mtllib import MaterialPool                                             # noqa: F401
    from ptina_b200.image import ImagePool                                                 # noqa: F401
    from ptina_b200.light import LightPool                                                 # noqa: F401
    from ptina_b200.light.world import WorldLight                                          # noqa: F401
    from ptina_b200.camera import Camera                                                   # noqa: F401
    from ptina_b200.filmtable import FilmTable                                             # noqa: F401
    from ptina_b200.sampling.sobol import SobolSampler                                     # noqa: F401


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, 'ptina_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(import|from)\s+oracle\b', text, flags=re.M), f'{f} imports the oracle'
                assert 'ptina_oracle' not in text and 'oracle/' not in text, f'{f} references the oracle'


def test_allocator_semantics():
    from ptina_b200.allocator import MemoryAllocator, IdAllocator
    m = MemoryAllocator(100)
    a, b = m.malloc(30), m.malloc(70)
    assert (a, b) == (0, 30)
    with pytest.raises(RuntimeError, match='Out of memory!'):
        m.malloc(1)
    m.free(a)
    assert m.malloc(10) == 0
    with pytest.raises(RuntimeError, match='Invalid pointer'):
        m.free(12345)
    ids = IdAllocator(2)
    assert (ids.malloc(), ids.malloc()) == (0, 1)
    with pytest.raises(RuntimeError, match='Out of ID!'):
        ids.malloc()


def test_material_factor_normalisation():
    import numpy as np
    from ptina_b200.mtllib import _expand_factor
    assert _expand_factor(None) == [1.0] * 4
    assert _expand_factor(0.5) == [0.5] * 4
    assert _expand_factor([0.1, 0.2, 0.3]) == [0.1, 0.2, 0.3, 1.0]
    assert _expand_factor(np.array(0.25)) == [0.25] * 4
    assert _expand_factor(np.array([1, 2, 3, 4.0])) == [1.0, 2.0, 3.0, 4.0]


def test_multimesh_contract():
    import numpy as np
    from ptina_b200.multimesh import compose_multiple_meshes
    from ptina_b200.tools import matrix as mx
    p = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]]], dtype=np.float32)
    n = np.array([[[0, 0, 1]] * 3], dtype=np.float32)
    v, m = compose_multiple_meshes([(p, n, None, mx.translate((1, 2, 3)) @ mx.scale(2), 5), (p, n, None, np.eye(4), None)])
    assert v.shape == (6, 8) and m.tolist() == [5, -1]
    assert np.allclose(v[1, :3], [3, 2, 3]) and np.allclose(v[0, 3:6], [0, 0, 1]) and np.allclose(v[:, 6:], 0)


def test_readobj_roundtrip(tmp_path):
    import numpy as np
    from ptina_b200.tools.readobj import readobj
    p = tmp_path / 'quad.obj'
    p.write_text('v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\nf 1/1/1 2/2/1 3/3/1 4/4/1\n')
    o = readobj(str(p))
    assert o['f'].shape == (2, 3, 3) and o['v'].shape == (4, 3)
    assert o['f'][1, :, 0].tolist() == [0, 2, 3] and o['f'][0, :, 1].tolist() == [0, 1, 2]
    verts = o['v'][o['f'][:, :, 0]].reshape(-1, 3)
    assert np.allclose(verts[3], [0, 0, 0]) and np.allclose(verts[5], [0, 1, 0])
