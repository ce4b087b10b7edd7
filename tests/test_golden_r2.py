"""Round-2 golden vectors (tests/golden/make_golden_r2.py: the reference's own sources under the shim): erfinv / normaldist
(common.py:337-352), PreviewEngine AOV passes (engine/preview.py:19-41) and render_tile (engine/path.py:96-118, compiled from the
reference's disabled text).  The oracle is held to them bit for bit; the CUDA path bit for bit where only +,-,*,/,sqrt are
involved (preview passes) and to 1e-5 / the image gate where libm or discrete shading decisions are."""
import os

import numpy as np
import pytest

import oracle
from ptina_b200 import scenes

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def gold(name):
    return np.load(os.path.join(G, name + '.npz'))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


def _scene(g, prefix, name):
    sc = scenes.CONFIGS[name]()
    sc['size'] = tuple(int(x) for x in g[prefix + 'size'])
    assert np.array_equal(np.asarray(sc['vertices'], np.float64), g[prefix + 'vertices']) and np.array_equal(sc['mtlids'], g[prefix + 'mtlids']), \
        'scene generator drifted from the golden fixture: regenerate tests/golden (make_golden_r2.py)'
    return sc


# ---- S6 ------------------------------------------------------------------------------------------------------------------------------
def test_normaldist_oracle_bit_exact():
    g = gold('normaldist')
    assert np.array_equal(bits(oracle.normaldist(g['samp'])), bits(g['normaldist']))
    assert np.array_equal(bits(oracle.erfinv(g['erfinv_x'])), bits(g['erfinv']))
    # sanity of the fixture itself: Winitzki's approximation is within 2e-3 of the true inverse error function
    from scipy.special import erfinv as true_erfinv
    x = g['erfinv_x'].astype(np.float64)
    inner = np.abs(x) < 0.999
    assert np.abs(g['erfinv'][inner] - true_erfinv(x[inner])).max() < 5e-3


@pytest.mark.gpu
def test_normaldist_gpu(gpu):
    g = gold('normaldist')
    got = gpu.normaldist(g['samp'])
    want = g['normaldist']
    # logf differs by <= 2 ulp between CUDA and glibc; near samp = 0.5 the result passes through sqrt(-tt1 + sqrt(tt1^2 - tt2)), a
    # cancellation that amplifies it (|normaldist| < 1e-3 there): absolute tolerance for those, 1e-5 relative for the rest
    big = np.abs(want) > 1e-2
    assert (np.abs(got[big] - want[big]) / np.abs(want[big])).max() <= 1e-5
    assert np.abs(got[~big] - want[~big]).max() <= 1e-5
    assert np.array_equal(np.sign(got), np.sign(want))


# ---- f1: PreviewEngine -----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('name', ['cornell_monkey', 'mini_matball'])
def test_preview_oracle_bit_exact(name):
    g = gold('preview')
    sc = _scene(g, name + '_', name)
    o = oracle.Oracle()
    scenes.apply(o, sc)
    ks = g[name + '_ks'].tolist()
    o.render_preview(len(ks), k_first=ks[0])
    for p in (1, 2):
        assert np.array_equal(bits(o.get_film(p)), bits(g[f'{name}_film{p}'])), f'{name}: film pass {p}'
        assert np.array_equal(bits(o.get_image(p)), bits(g[f'{name}_image{p}']))
    assert not o.get_film(0).any() and not g[name + '_film0'].any()           # pass 0 untouched (preview.py:40-41)


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['cornell_monkey', 'mini_matball'])
def test_preview_gpu_matches_reference_golden(gpu, name):
    """Albedo (MaterialPool.get incl. the bilinear texture fetch) and the interpolated, ray-facing shading normal of the primary hit:
    +,-,*,/,sqrt only, so the CUDA passes equal the reference's bit for bit."""
    from ptina_b200 import worker
    from ptina_b200.engine import PreviewEngine
    g = gold('preview')
    sc = _scene(g, name + '_', name)
    gpu.sobol_reset()
    scenes.apply(worker, sc)
    worker.clear()
    ks = g[name + '_ks'].tolist()
    PreviewEngine().render_range(ks[0], len(ks), 1)
    for p in (1, 2):
        got, want = gpu.get_film(p), g[f'{name}_film{p}']
        assert np.array_equal(got[..., 3], want[..., 3])
        assert np.array_equal(bits(got), bits(want)), f'{name}: pass {p} differs in {(bits(got) != bits(want)).sum()} words, max {np.abs(got - want).max()}'
        assert np.array_equal(bits(worker.get_image(p)), bits(g[f'{name}_image{p}']))
    assert not gpu.get_film(0).any()
    # and through the reference's call pattern: PreviewEngine().render() advances the Sobol time like every engine
    worker.clear(); gpu.sobol_time = ks[0] - 1
    for _ in ks:
        worker.render_preview()
    assert gpu.sobol_time == ks[-1]
    assert np.array_equal(bits(gpu.get_film(1)), bits(g[f'{name}_film1']))


# ---- f4: render_tile ---------------------------------------------------------------------------------------------------------------------
def _tile_reference(g):
    """In-range part of what the reference text renders: pixels (x >= 1, y == 0) also received the two samples the text's `y > ny`
    check lets through at y == ny (see make_golden_r2.py); they are compared separately."""
    text = g['film_text']
    nx, ny = text.shape[:2]
    clean = np.ones((nx, ny), bool)
    clean[1:, 0] = False
    return text, clean


def test_tile_oracle_matches_reference_text():
    g = gold('tile')
    sc = _scene(g, '', 'cornell_monkey')
    o = oracle.Oracle()
    scenes.apply(o, sc)
    o.sobol_time = int(g['k_before'])
    o.render_tile(oracle.ENGINE_PATH, 0, 0, int(g['samples']))
    assert o.sobol_time == int(g['k_tile'])
    text, clean = _tile_reference(g)
    film = o.get_film(0)
    assert (film[..., 3] == 2).all() and (text[..., 3][clean] == 2).all() and (text[..., 3][~clean] == 4).all()
    assert np.array_equal(bits(film[clean]), bits(text[clean]))
    # aliased pixels: the text's extra contributions came first in its loop order, then the pixel's own two samples
    own = film[~clean][:, :3]
    extra = g['alias'][~clean][:, :3]
    assert np.allclose(text[~clean][:, :3], extra + own, rtol=1e-6, atol=1e-7)
    # render_final = render_tile over every tile, 64 samples at a time (path.py:120-128)
    o.clear(); o.sobol_time = 64
    o.render_final(oracle.ENGINE_PATH, 70)
    f = o.get_film(0)
    assert (f[..., 3] == 64 + 7).all() and o.sobol_time == 64 + 2          # one tile, two render_tile calls: 64 samples, then 0..6


@pytest.mark.gpu
def test_tile_gpu(gpu):
    from ptina_b200 import worker, _native
    from ptina_b200.engine import PathEngine
    g = gold('tile')
    sc = _scene(g, '', 'cornell_monkey')
    gpu.sobol_reset()
    scenes.apply(worker, sc)
    worker.clear()
    gpu.sobol_time = int(g['k_before'])
    PathEngine().render_tile(0, 0, int(g['samples']))
    assert gpu.sobol_time == int(g['k_tile'])
    text, clean = _tile_reference(g)
    film = gpu.get_film(0)
    assert (film[..., 3] == 2).all()
    rel = np.abs(film[..., :3] - text[..., :3]).max(2)[clean] / np.maximum(np.abs(text[..., :3]).max(2)[clean], 1e-2)
    assert np.median(rel) < 1e-6 and (rel > 1e-4).mean() < 0.03, (float(np.median(rel)), float((rel > 1e-4).mean()))
    # full-size tiles against the oracle: a 130x70 film has 3x2 tiles, the last ones partial
    sc2 = dict(scenes.cornell_monkey(), size=(130, 70))
    scenes.apply(worker, sc2)
    worker.clear(); gpu.sobol_reset()
    PathEngine().render_final(3)                                             # samples = 3 -> m = 0..3: four samples per pixel
    o = oracle.Oracle(); scenes.apply(o, sc2)
    o.render_final(oracle.ENGINE_PATH, 3)
    a, b = gpu.get_film(0), o.get_film(0)
    assert gpu.sobol_time == o.sobol_time == 64 + 6 and (a[..., 3] == 4).all() and np.array_equal(a[..., 3], b[..., 3])
    rmse = np.sqrt((((a[..., :3] - b[..., :3]) / np.maximum(b[..., :3], 4e-2)) ** 2).mean())
    assert rmse < 1e-3, rmse
    # out-of-film tiles are a no-op apart from the Sobol update; the brute engine has tiles too
    t = gpu.sobol_time
    gpu.render_tile(_native.ENGINE_PATH, 5, 0, 3)
    assert gpu.sobol_time == t + 1 and np.array_equal(gpu.get_film(0), a)
    gpu.render_tile(_native.ENGINE_BRUTE, 0, 0, 0)
    assert gpu.get_film(0)[:64, :64, 3].min() == 5
