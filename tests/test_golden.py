"""Pins the oracle (and, under -m gpu, the CUDA path) to golden vectors produced by executing the REFERENCE'S OWN SOURCES
(/root/reference/ptina/*.py, unmodified) under the Taichi-semantics shim oracle/tishim -- see tests/golden/make_golden.py for
how they were made and what is not the reference's code in such a run.  Integer / index results: bit-exact.  f32 results of
+,-,*,/,sqrt chains (rays, boxes, hit depth/uv): bit-exact.  Results that pass through libm (sin/cos/atan2/log/pow): the
oracle links the same glibc the shim calls, so they are bit-exact too; the CUDA path uses CUDA's libm and is held to 1e-5."""
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


def scene_of(name, g):
    sc = scenes.CONFIGS[name]()
    sc['size'] = tuple(int(x) for x in g['size'])
    assert np.array_equal(np.asarray(sc['vertices'], np.float64), g['vertices']) and np.array_equal(sc['mtlids'], g['mtlids']), \
        'scene generator drifted from the golden fixture: regenerate tests/golden (make_golden.py)'
    return sc


def load_oracle(name, g):
    sc = scene_of(name, g)
    o = oracle.Oracle()
    scenes.apply(o, sc)
    if 'extra_light_world' in g:
        o.add_light(g['extra_light_world'], g['extra_light_color'], float(g['extra_light_size']), 'POINT')
    return sc, o


def test_sobol_table_and_state():
    g = gold('sobol')
    from oracle.sobol_table import vgrid_i32
    V = vgrid_i32()
    assert np.array_equal(V[1:], g['V'][1:]), 'direction numbers differ from calc_sobol_vgrid run from the reference source'
    assert int(g['time']) == 64
    o = oracle.Oracle()
    assert np.array_equal(bits(o.sobol_point(64)), bits(g['P64']))           # state after SobolSampler.reset() (64 updates)
    from ptina_b200.sampling.sobol import calc_sobol_vgrid
    assert np.array_equal(calc_sobol_vgrid()[1:], g['V'][1:])                 # the product's own table builder


def test_integer_functions():
    g = gold('integer')
    for (a, b), want in zip(g['wh_ij'], g['wh']):
        assert oracle.wanghash2(int(a), int(b)) == int(want)
    for p, want in zip(g['morton_pts'], g['morton']):
        assert oracle.morton3d(*[float(x) for x in p]) == int(want)
    for x, want in zip(g['clz_x'], g['clz']):
        assert oracle.clz(int(x)) == int(want)
    from ptina_b200.sampling import wanghash2
    assert all(wanghash2(int(a), int(b)) == int(w) for (a, b), w in zip(g['wh_ij'][:64], g['wh'][:64]))


def _bsdf_inputs(g):
    m = g['params'].shape[0]
    geom_e = np.concatenate([g['normal'], g['sign'][:, None], g['wi'], g['wo']], 1).astype(np.float32)
    geom_s = np.concatenate([g['normal'], g['sign'][:, None], g['wi'], g['samp']], 1).astype(np.float32)
    return m, geom_e, geom_s


def test_disney_bsdf_bit_exact():
    g = gold('bsdf')
    m, geom_e, geom_s = _bsdf_inputs(g)
    ev = oracle.eval_bsdf(g['params'], geom_e)
    sm = oracle.sample_bsdf(g['params'], geom_s)
    # eval: +,-,*,/,sqrt and logf (GTR1) only
    nan_e = np.isnan(g['eval'])
    assert np.array_equal(np.isnan(ev), nan_e)
    assert np.array_equal(bits(ev)[~nan_e], bits(g['eval'])[~nan_e]), f'{(bits(ev) != bits(g["eval"])).sum()} eval words differ'
    nan_s = np.isnan(g['sample'])
    assert np.array_equal(np.isnan(sm), nan_s)
    assert np.array_equal(bits(sm)[~nan_s], bits(g['sample'])[~nan_s]), f'{(bits(sm)[~nan_s] != bits(g["sample"])[~nan_s]).sum()} sample words differ'
    # all lobes present in the fixture: invalid (0 pdf), diffuse (1/pi), specular, reflection/refraction through glass, coat
    pdf = g['sample'][:, 3]
    assert (pdf == 0).any() and np.isclose(pdf, 1 / np.pi).any() and ((g['sample'][:, :3] * g['normal']).sum(1) < -0.1).any()


@pytest.mark.parametrize('name', ['cornell_boxes', 'cornell_monkey', 'mini_matball'])
def test_tree_rays_hits_bit_exact(name):
    g = gold(name)
    sc, o = load_oracle(name, g)
    t = o.export_tree()
    for key in ('mc', 'id', 'leaf', 'child'):
        assert np.array_equal(t[key], g['tree_' + key]), f'{name}: {key}'
    assert np.array_equal(bits(t['bmin']), bits(g['tree_bmin'])) and np.array_equal(bits(t['bmax']), bits(g['tree_bmax']))
    p = o.primary(int(g['k_primary']))
    assert np.array_equal(bits(p['rays']), bits(g['rays'])), f'{name}: primary rays'
    h = g['hits']
    assert np.array_equal(p['hit'], h[:, 0].astype(np.int32)) and np.array_equal(p['index'], h[:, 2].astype(np.int32))
    assert np.array_equal(bits(p['depth']), bits(h[:, 1])) and np.array_equal(bits(p['uv']), bits(h[:, 3:5]))
    assert (p['hit'] == 1).sum() > 20


@pytest.mark.parametrize('name,engine', [('cornell_boxes', oracle.ENGINE_PATH), ('cornell_monkey', oracle.ENGINE_PATH), ('mini_matball', oracle.ENGINE_BRUTE)])
def test_radiance_and_film_bit_exact(name, engine):
    g = gold(name)
    sc, o = load_oracle(name, g)
    rad = o.render_sample(engine, int(g['k_primary']))
    assert np.array_equal(bits(rad), bits(g['radiance'])), f'{name}: {(bits(rad) != bits(g["radiance"])).sum()} radiance words differ'
    ks = g['frame_ks'].tolist()
    assert ks == list(range(ks[0], ks[0] + len(ks)))
    o.render(engine, len(ks), k_first=ks[0])
    assert np.array_equal(bits(o.get_film()), bits(g['film']))
    assert np.array_equal(bits(o.get_image()), bits(g['image']))
    out = np.zeros(g['fast_export'].shape, np.float32)
    o.fast_export_image(out)
    assert np.array_equal(bits(out), bits(g['fast_export']))


def test_lights_bit_exact():
    for name in ('cornell_boxes', 'cornell_monkey'):
        g = gold(name)
        sc, o = load_oracle(name, g)
        lh = o.light_hit(np.concatenate([g['light_org'], g['light_dir']], 1))
        assert np.array_equal(bits(lh), bits(g['light_hit']))
        ls = o.light_sample(np.concatenate([g['light_org'], g['light_samp']], 1))
        assert np.array_equal(bits(ls), bits(g['light_sample']))
        assert (g['light_hit'][:, 0] == 1).any()


def test_materials_textures_environment_bit_exact():
    g = gold('mini_matball')
    sc, o = load_oracle('mini_matball', g)
    mg = o.material_get(g['mat_id'], g['mat_uv'])
    assert np.array_equal(bits(mg), bits(g['mat_get']))
    wa = o.world_at(g['world_dir'])
    assert np.array_equal(bits(wa), bits(g['world_at']))


# ------------------------------------------------------------------------------------------------------------------------------
# the same fixtures against the CUDA path
# ------------------------------------------------------------------------------------------------------------------------------
def _load_gpu(gpu, name, g):
    from ptina_b200 import worker
    sc = scene_of(name, g)
    gpu.sobol_reset()
    scenes.apply(worker, sc)
    if 'extra_light_world' in g:
        worker.add_light(g['extra_light_world'], g['extra_light_color'], float(g['extra_light_size']), 'POINT')
    worker.clear()
    return sc


@pytest.mark.gpu
@pytest.mark.parametrize('name', ['cornell_boxes', 'cornell_monkey', 'mini_matball'])
def test_gpu_against_reference_golden(gpu, name):
    from ptina_b200 import _native
    g = gold(name)
    _load_gpu(gpu, name, g)
    t = gpu.export_tree()
    for key in ('mc', 'id', 'leaf', 'child'):
        assert np.array_equal(t[key], g['tree_' + key]), f'{name}: {key}'
    assert np.array_equal(bits(t['bmin']), bits(g['tree_bmin'])) and np.array_equal(bits(t['bmax']), bits(g['tree_bmax']))
    p = gpu.trace_primary(int(g['k_primary']))
    h = g['hits']
    assert np.array_equal(bits(p['rays']), bits(g['rays']))
    assert np.array_equal(p['index'], h[:, 2].astype(np.int32)) and np.array_equal(p['hit'], h[:, 0].astype(np.int32))
    hit = h[:, 0] == 1
    assert np.array_equal(bits(p['depth'])[hit], bits(h[:, 1])[hit]) and np.array_equal(bits(p['uv'])[hit], bits(h[:, 3:5])[hit])
    eng = _native.ENGINE_BRUTE if name == 'mini_matball' else _native.ENGINE_PATH
    rad = gpu.render_sample(eng, int(g['k_primary']))
    rel = np.abs(rad - g['radiance']).max(2) / np.maximum(np.abs(g['radiance']).max(2), 1e-2)
    assert np.median(rel) < 1e-6 and (rel > 1e-4).mean() < 0.02, (float(np.median(rel)), float((rel > 1e-4).mean()))
    ks = g['frame_ks'].tolist()
    gpu.render_range(eng, ks[0], len(ks), 1)
    img = gpu.get_image()
    rmse = np.sqrt((((img[..., :3] - g['image'][..., :3]) / np.maximum(g['image'][..., :3], 1e-2)) ** 2).mean())
    assert rmse < 1e-3, rmse


@pytest.mark.gpu
def test_gpu_bsdf_and_taps_against_reference_golden(gpu):
    g = gold('bsdf')
    m, geom_e, geom_s = _bsdf_inputs(g)
    ev = gpu.eval_bsdf(g['params'], geom_e)
    ok = ~np.isnan(g['eval']).any(1)
    scale = np.maximum(np.abs(g['eval'][ok]).max(1, keepdims=True), 1e-4)
    assert (np.abs(ev[ok] - g['eval'][ok]) / scale).max() <= 1e-5
    sm = gpu.sample_bsdf(g['params'], geom_s)
    ok = ~np.isnan(g['sample']).any(1)
    assert np.array_equal(np.isnan(sm).any(1), ~ok)
    assert np.abs(sm[ok, :3] - g['sample'][ok, :3]).max() <= 2e-5
    g2 = gold('mini_matball')
    _load_gpu(gpu, 'mini_matball', g2)
    mg = gpu.material_get(g2['mat_id'], g2['mat_uv'])
    assert (np.abs(mg - g2['mat_get']) / np.maximum(np.abs(g2['mat_get']), 1e-3)).max() <= 1e-5
    g3 = gold('cornell_boxes')
    _load_gpu(gpu, 'cornell_boxes', g3)
    lh = gpu.light_hit(np.concatenate([g3['light_org'], g3['light_dir']], 1))
    assert np.array_equal(lh[:, 0], g3['light_hit'][:, 0])
    assert (np.abs(lh[:, 1:] - g3['light_hit'][:, 1:]) / np.maximum(np.abs(g3['light_hit'][:, 1:]), 1e-4)).max() <= 1e-5
