"""Oracle-anchored gates at BASELINE.json's stated sizes.

  * configs 1 and 2, 512x512, 1024 spp: north_star's image gate -- per-pixel relative RMSE <= 1e-3 against the converged image the
    CPU oracle rendered from the same Sobol indices (tests/golden/image_*.npz, made by tests/golden/make_oracle_images.py; the
    oracle itself is pinned bit for bit to the reference's sources by tests/test_golden.py).
  * config 3, 1024x1024, BruteEngine with transmission, textures and the environment map: the same comparison at 64 spp (a converged
    1024^2 x 256 spp oracle image costs CPU-hours); with so few samples a single libm-induced decision flip moves a pixel by 1/64
    of a sample, so the gate is the per-pixel one the per-sample tests use (median, fraction beyond 1e-4) and the RMSE is reported.
  * config 4, 1 024 010 triangles: the oracle builds the same tree on the host (bit-exact arrays) and traces primary rays across the
    whole 1920x1080 frame and incoherent secondary rays with `avoid`; the production traversal returns the same hit ids, depths and
    barycentrics bit for bit, and the same shadow verdicts.
  * the reference's per-call pattern: 32 x PathEngine().render() records samples that are submitted as one batch -- same film bits.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from ptina_b200 import scenes, worker, _native  # noqa: E402

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


def rel_rmse(img, ref, floor=1e-2):
    """Per-pixel relative RMSE over rgb: differences relative to the reference value (floored, so black pixels do not divide by 0)."""
    return float(np.sqrt((((img - ref) / np.maximum(ref, floor)) ** 2).mean()))


def _render_like_golden(gpu, name):
    g = np.load(os.path.join(G, f'image_{name}.npz'))
    sc = scenes.CONFIGS[name]()
    assert tuple(g['size']) == tuple(sc['size'])
    gpu.sobol_reset()
    scenes.apply(worker, sc)
    worker.clear()
    eng = {'path': _native.ENGINE_PATH, 'brute': _native.ENGINE_BRUTE}[sc['engine']]
    gpu.set_counting(True, False); gpu.reset_counters()
    gpu.render_range(eng, int(g['k_first']), int(g['spp']), 1)
    img = gpu.get_image()
    cnt = gpu.counters()
    gpu.set_counting(False, False)
    st = int(g['stride'])
    assert (gpu.get_film()[..., 3] == int(g['spp'])).all()
    return g, img[::st, ::st, :3], cnt


@pytest.mark.parametrize('name', ['cornell_boxes', 'cornell_monkey'])
def test_converged_image_gate_1024spp(gpu, name):
    g, img, cnt = _render_like_golden(gpu, name)
    ref = g['rgb']
    rmse = rel_rmse(img, ref)
    off = np.abs(img - ref).max(2) / np.maximum(np.abs(ref).max(2), 1e-2)
    print(f'{name}: 512x512 x {int(g["spp"])} spp rel-RMSE {rmse:.3e}, pixels off by > 1e-3: {(off > 1e-3).sum()}, max {off.max():.2e}')
    assert rmse <= 1e-3, rmse
    assert (off > 1e-3).mean() < 1e-3
    # the oracle counts every shadow ray, the device skips those whose contribution is exactly zero: never more rays than the oracle
    assert 0.9 * int(g['rays']) <= cnt['rays'] <= 1.001 * int(g['rays'])


def test_fast_mode_holds_the_image_gate_not_the_tap_gate(gpu):
    """PTB_MODE_FAST (opt-in, non-parity): the shading stage's FMA-contracted / approximate-division build.  Trees, rays and hit ids are
    untouched; the converged-image gate still holds; BSDF taps move by more than the 1e-5 the parity mode keeps -- which is why it is
    not the default."""
    gpu.set_mode('fast')
    try:
        assert gpu.mode == _native.MODE_FAST
        g, img, _ = _render_like_golden(gpu, 'cornell_monkey')
        rmse = rel_rmse(img, g['rgb'])
        prim_fast = gpu.trace_primary(66)
        bs = np.load(os.path.join(G, 'bsdf.npz'))
        geom = np.concatenate([bs['normal'], bs['sign'][:, None], bs['wi'], bs['wo']], 1).astype(np.float32)
        ev_fast = gpu.eval_bsdf(bs['params'], geom)
    finally:
        gpu.set_mode('parity')
    assert gpu.mode == _native.MODE_PARITY
    prim = gpu.trace_primary(66)
    assert np.array_equal(prim['index'], prim_fast['index']) and np.array_equal(bits(prim['depth']), bits(prim_fast['depth']))
    ev = gpu.eval_bsdf(bs['params'], geom)
    ok = ~np.isnan(ev).any(1)
    tap = float((np.abs(ev_fast[ok] - ev[ok]) / np.maximum(np.abs(ev[ok]).max(1, keepdims=True), 1e-4)).max())
    print(f'fast mode: 1024-spp rel-RMSE {rmse:.3e}; BSDF eval vs parity mode: max rel {tap:.2e}')
    assert rmse <= 1e-3, rmse
    assert tap < 5e-3


def test_config3_full_resolution(gpu):
    g, img, _ = _render_like_golden(gpu, 'matball')
    ref = g['rgb']
    rel = np.abs(img - ref).max(2) / np.maximum(np.abs(ref).max(2), 1e-2)
    rmse = rel_rmse(img, ref)
    print(f'matball: 1024x1024 x {int(g["spp"])} spp median rel {np.median(rel):.2e}, beyond 1e-4: {(rel > 1e-4).mean():.4f}, rel-RMSE {rmse:.3e}')
    assert np.median(rel) < 1e-6 and (rel > 1e-4).mean() < 0.02
    assert rmse < 2e-2
    # bit-exact part at full size: primary hits of a Sobol point against the oracle
    sc = scenes.matball()
    o = oracle.Oracle(); scenes.apply(o, sc)
    a, b = gpu.trace_primary(70), o.primary(70)
    assert np.array_equal(bits(a['rays']), bits(b['rays'])) and np.array_equal(a['index'], b['index'])
    h = b['hit'] == 1
    assert np.array_equal(bits(a['depth'])[h], bits(b['depth'])[h]) and np.array_equal(bits(a['uv'])[h], bits(b['uv'])[h])


def test_config4_full_size_against_oracle(gpu):
    sc = scenes.mega()
    gpu.sobol_reset()
    scenes.apply(worker, sc)
    worker.clear()
    o = oracle.Oracle(); scenes.apply(o, sc)
    assert gpu.tree.n == len(sc['mtlids']) > 1_000_000 and gpu.tree.policy == _native.TRAVERSE_ORDERED
    assert gpu.tree.trav_ploc == 1 and 0 < gpu.tree.trav_depth <= 64       # the multi-block PLOC tree is what the production policy walks
    ta, tb = gpu.export_tree(), o.export_tree()
    for key in ('mc', 'id', 'leaf', 'child'):
        assert np.array_equal(ta[key], tb[key]), key
    assert np.array_equal(bits(ta['bmin']), bits(tb['bmin'])) and np.array_equal(bits(ta['bmax']), bits(tb['bmax']))
    assert o.validate_tree() == gpu.tree.depth
    # every primary ray of the 1920x1080 frame: production policy vs the oracle's unordered DFS
    k = 65
    a, b = gpu.trace_primary(k), o.primary(k)
    assert np.array_equal(bits(a['rays']), bits(b['rays']))
    assert np.array_equal(a['index'], b['index']) and np.array_equal(a['hit'], b['hit'])
    h = b['hit'] == 1
    assert h.mean() > 0.5
    assert np.array_equal(bits(a['depth'])[h], bits(b['depth'])[h]) and np.array_equal(bits(a['uv'])[h], bits(b['uv'])[h])
    rng = np.random.default_rng(11)
    # incoherent secondary rays from random surface points, with avoid = the triangle they leave
    m = 65536
    verts = np.asarray(sc['vertices'], np.float32)[:, :3].reshape(-1, 3, 3)
    f = rng.integers(0, verts.shape[0], m)
    w = rng.dirichlet([1, 1, 1], m).astype(np.float32)
    org = (verts[f] * w[:, :, None]).sum(1)
    d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([org, d], 1).astype(np.float32)
    avoid = f.astype(np.int32)
    ref = o.intersect(rays, avoid)
    got = gpu.intersect(rays, avoid, _native.TRAVERSE_ORDERED)
    assert np.array_equal(got['index'], ref['index']), f'{(got["index"] != ref["index"]).sum()} hit ids differ'
    h = ref['hit'] == 1
    assert np.array_equal(bits(got['depth'])[h], bits(ref['depth'])[h]) and np.array_equal(bits(got['uv'])[h], bits(ref['uv'])[h])
    dis = np.where(h, ref['depth'] * rng.choice([0.5, 1.0, 1.5], m), 5.0).astype(np.float32)
    want = (h & (ref['depth'] <= dis)).astype(np.int32)
    assert np.array_equal(gpu.occluded(rays, dis, avoid, _native.TRAVERSE_ORDERED), want)
    # one path-traced sample of a 240x135 crop-sized film of the same scene against the oracle (radiance, per pixel)
    worker.set_size(240, 135); o.set_size(240, 135)
    ra, rb = gpu.render_sample(_native.ENGINE_PATH, 66), o.render_sample(oracle.ENGINE_PATH, 66)
    rel = np.abs(ra - rb).max(2) / np.maximum(np.abs(rb).max(2), 1e-2)
    assert np.median(rel) < 1e-6 and (rel > 1e-4).mean() < 0.02, (float(np.median(rel)), float((rel > 1e-4).mean()))


def test_two_lane_schedules_are_bit_identical(gpu):
    """Multi-batch renders alternate half-sized chunks between two lanes, the big-tree PLOC rebuild changes the traversal tree, and trees
    walked out of global memory use 4-wide nodes: none of it may change a bit of the film (PTB_NO_RESIDENT_BVH=1 runs of the suite put
    every scene through the global-memory kernel).  1920x1080 fills the path pool with 4 samples, so 10 samples are 5 chunks on 2 lanes."""
    sc = dict(scenes.mega_small(), size=(1920, 1080))
    films = {}
    for key, opts in (('two', {}), ('one', {'pt_lanes': 1}), ('four', {'pt_lanes': 4}), ('lbvh', {'ploc_big': 0}), ('binary', {'wide4': 0})):
        for k, v in opts.items():
            gpu.set_option(k, v)
        try:
            gpu.sobol_reset()
            scenes.apply(worker, sc)
            worker.clear()
            gpu.render(_native.ENGINE_PATH, 10)
            films[key] = gpu.get_film().copy()
            assert gpu.tree.trav_ploc == (0 if key == 'lbvh' else 1)
        finally:
            for k in opts:
                gpu.set_option(k, {'pt_lanes': 2}.get(k, 1))
    assert (films['two'][..., 3] == 10).all()
    for key in ('one', 'four', 'lbvh', 'binary'):
        assert np.array_equal(bits(films['two']), bits(films[key])), key
    scenes.apply(worker, sc)       # back to the default tree


def test_percall_render_is_coalesced(gpu):
    """exams/benchmark.py:29-33 calls PathEngine().render() once per sample.  The library records those calls and submits them as one
    wavefront batch at the first readback: same Sobol points, same per-pixel accumulation order -> the same film bits as one
    render(32), and far fewer kernel launches than 32 one-sample wavefronts."""
    from ptina_b200.engine import PathEngine
    sc = scenes.cornell_monkey()
    gpu.sobol_reset()
    scenes.apply(worker, sc)
    worker.clear()
    gpu.reset_counters()
    PathEngine().render(32)
    a = gpu.get_film().copy()
    la = gpu.launches()
    worker.clear(); gpu.sobol_reset(); gpu.reset_counters()
    for i in range(32):
        if i == 31:
            assert gpu.launches() == 0                # 31 recorded calls: nothing has been submitted yet
        PathEngine().render()
        assert gpu.sobol_time == 65 + i               # the generator's time advances call by call, as path.py:75-77 does
    b = gpu.get_film().copy()
    lb = gpu.launches()
    assert np.array_equal(bits(a), bits(b)) and (a[..., 3] == 32).all()
    assert lb <= la + 4, (la, lb)                      # one batch (the 32nd call fills it), not 32 one-sample wavefronts
    # anything that could observe the difference flushes first: a size change renders the recorded samples at the OLD size
    worker.clear(); gpu.sobol_reset()
    PathEngine().render(); PathEngine().render()
    worker.set_size(64, 64)
    assert gpu.sobol_time == 66
    worker.set_size(*sc['size'])
    c = gpu.get_film()
    assert (c[..., 3] == 2).all()
    # engines are not merged with each other, and an interleaved sequence keeps its order
    worker.clear(); gpu.sobol_reset()
    from ptina_b200.engine import PreviewEngine
    PathEngine().render(); PreviewEngine().render(); PathEngine().render()
    assert (gpu.get_film(0)[..., 3] == 2).all() and (gpu.get_film(1)[..., 3] == 1).all() and gpu.sobol_time == 67


def _fresh_process_state(gpu):
    """What a new process starts with (the scripts never touch these): zeroed material tables and no images, SobolSampler.reset()
    (64 updates), LightPool's default point light (light/__init__.py:22-28), WorldLight's 0.1 factor with the zero-initialised texture id (light/world.py:10-16)."""
    from ptina_b200.tools import matrix as mx
    from ptina_b200.mtllib import MaterialPool
    MaterialPool().fac[:] = 0; MaterialPool().tex[:] = 0      # the zero-initialised tables a short (glTF, 3-slot) material list leaves alone
    worker.load_images([])
    gpu.sobol_reset()
    worker.clear_lights()
    worker.add_light(mx.translate((1.0, 2.0, 3.0)), np.array([32.0, 32.0, 32.0]), 0.5, 'POINT')
    worker.set_world_light([0.1] * 4, 0)


def test_unmodified_benchmark_script(gpu, tmp_path, capsys, monkeypatch):
    """The reference's exams/benchmark.py -- its text, unmodified (tests/golden/exams/benchmark.py.txt; tests/test_compat_cpu.py checks
    the copy against /root/reference where that exists) -- runs against the facade through the `ptina` / `taichi` compatibility
    packages and the generated assets/monkey_cornell.gltf, and leaves the film its API calls describe."""
    from ptina_b200 import compat, things
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / 'benchmark.py'
    script.write_text(open(os.path.join(G, 'exams', 'benchmark.py.txt')).read())
    os.symlink(os.path.join(root, 'assets'), tmp_path / 'assets')
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(compat.PATH)
    _fresh_process_state(gpu)
    ns = compat.run_script(str(script))
    out = capsys.readouterr().out
    assert '31 samples...' in out and out.strip().endswith('sps')
    import taichi as ti
    assert ti.shown[-1][1] == (512, 512, 4) and ti.init_args['arch'] == 'cuda'
    img = ns['img']
    assert img.shape == (512, 512, 4) and (img[..., 3] == 1).all()
    # the same calls made directly: glTF ingest -> 3-slot materials, default point light, Sobol points 66..97 after the warm-up frame
    from ptina_b200.tools.readgltf import readgltf
    from ptina_b200.engine import PathEngine
    v, m, mats, imgs = readgltf(os.path.join(root, 'assets', 'monkey_cornell.gltf'))
    assert gpu.sobol_time == 64 + 33 and gpu.nfaces == 978
    worker.clear()
    gpu.render_range(_native.ENGINE_PATH, 66, 32, 1)
    assert np.array_equal(bits(worker.get_image()), bits(img))
    for mod in [k for k in list(__import__('sys').modules) if k == 'ptina' or k.startswith('ptina.') or k == 'taichi']:
        del __import__('sys').modules[mod]


def test_unmodified_benchtiles_script(gpu, tmp_path, capsys, monkeypatch):
    """exams/benchtiles.py, unmodified: render_tile(0, 0, 0) as a warm-up, then render_final(32) -- 8x8 tiles of 64x64 pixels, samples
    m = 0..32 each (`m > samples: continue` is inclusive) -- against the oracle's restatement of the same loops."""
    from ptina_b200 import compat
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / 'benchtiles.py'
    script.write_text(open(os.path.join(G, 'exams', 'benchtiles.py.txt')).read())
    os.symlink(os.path.join(root, 'assets'), tmp_path / 'assets')
    monkeypatch.chdir(tmp_path)
    monkeypatch.syspath_prepend(compat.PATH)
    _fresh_process_state(gpu)
    ns = compat.run_script(str(script))
    assert capsys.readouterr().out.strip().endswith('sps')
    img = ns['img']
    film = gpu.get_film(0)
    assert img.shape == (512, 512, 4) and (film[..., 3] == 33).all() and gpu.sobol_time == 64 + 1 + 64
    # the oracle on one tile of the same frame (tile (3, 2) is the 3*8 + 2 + 1 = 27th render_tile call after the warm-up one)
    from ptina_b200.tools.readgltf import readgltf
    v, m, mats, imgs = readgltf(os.path.join(root, 'assets', 'monkey_cornell.gltf'))
    o = oracle.Oracle()
    o.load_materials(mats); o.load_images(imgs); o.load_model(v, m); o.build_tree()
    o.set_camera(scenes.BENCHMARK_PERS); o.set_size(512, 512)
    o.sobol_time = 64 + 1 + 3 * 8 + 2
    o.render_tile(oracle.ENGINE_PATH, 3, 2, 32)
    a, b = film[192:256, 128:192, :3] / 33, o.get_film(0)[192:256, 128:192, :3] / 33
    rel = np.abs(a - b).max(2) / np.maximum(np.abs(b).max(2), 1e-2)
    assert np.median(rel) < 1e-6 and rel_rmse(a, b) < 1e-3, (float(np.median(rel)), rel_rmse(a, b))
    for mod in [k for k in list(__import__('sys').modules) if k == 'ptina' or k.startswith('ptina.') or k == 'taichi']:
        del __import__('sys').modules[mod]


def test_input_validation(gpu):
    """Ids that would index past the material / texture tables are refused at load time (host arrays) or at build time (device
    arrays); taps refuse missing buffers; mlt_state refuses a chain count that is not the context's."""
    import torch
    from ptina_b200.model import ModelPool
    from ptina_b200.tree import BVHTree
    sc = scenes.cornell_boxes(nx=32, ny=32)
    scenes.apply(worker, sc)
    bad = np.array(sc['mtlids'], np.int32); bad[3] = 64
    with pytest.raises(_native.NativeError, match='material id 64'):
        ModelPool().load(sc['vertices'], bad)
    bad[3] = -2
    with pytest.raises(_native.NativeError, match='material id -2'):
        ModelPool().load(sc['vertices'], bad)
    dv = torch.from_numpy(np.ascontiguousarray(sc['vertices'], np.float32)).cuda()
    ModelPool().load(dv, torch.from_numpy(bad).cuda())
    with pytest.raises(RuntimeError, match='material ids outside'):
        BVHTree().build()
    with pytest.raises(AssertionError, match='float32'):
        ModelPool().load(dv.double(), torch.from_numpy(bad).cuda())
    with pytest.raises(_native.NativeError, match='texture id'):
        worker.set_world_light([1, 1, 1, 1], 64)
    with pytest.raises(_native.NativeError, match='texture id'):
        worker.load_materials([scenes.material(tex={'basecolor': 99})])
    scenes.apply(worker, sc)
    rays = np.zeros((4, 6), np.float32); rays[:, 5] = 1
    import ctypes
    assert gpu.L.ptb_occluded(gpu.h, rays.ctypes.data_as(ctypes.c_void_p), None, None, 4, 0, None) != 0
    assert b'needs rays, dis and occluded' in gpu.L.ptb_last_error()
    gpu.mlt_reset(0, 0, 1024)
    with pytest.raises(_native.NativeError, match='sized for 16 chains'):
        gpu.mlt_state(16)
    with pytest.raises(_native.NativeError, match='GL'):
        gpu.fast_export_gl(1)                 # no OpenGL context on this thread: a clean error, not a crash
    worker.clear()
