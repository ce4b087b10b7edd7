"""Generates tests/golden/*.npz by executing the REFERENCE'S OWN SOURCES (imported from /root/reference, unmodified) under
the Taichi-semantics shim in oracle/tishim.  Runs only in the authoring container (the GPU box has no /root/reference);
the .npz files it writes are committed.

    python tests/golden/make_golden.py [--only NAME]

What is NOT the reference's code in a golden run, and why:
  * `taichi` is oracle/tishim/taichi (strict-IEEE f32 / wrapping i32 semantics, sequential loops) and `pysobol` is a stub
    that serves SciPy's copy of the same Joe-Kuo table: neither package can be installed here.
  * ModelPool.from_numpy / ImagePool.from_numpy (pure host->field copies written with Taichi's write-through element
    proxies) are replaced by direct NumPy copies into the same fields.
  * Python's `a[b]` on the reference's `is_taichi_class` objects is routed to their own `subscript` methods (Taichi's AST
    transformer does this), and the builtins int/float/min/max/abs resolve to their Taichi meanings inside ptina modules.
  * np.argsort in sortMortonCodes is made stable (`kind='stable'`): its tie order is unspecified in the reference and the
    canonical choice of this project is the stable one.
Everything else -- every formula of the hot path -- is the reference's own text.
"""
import argparse
import contextlib
import io
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'tishim'))

import taichi as ti  # noqa: E402  (the shim)

assert 'tishim' in ti.__file__

# ---- import the reference package without running ptina/__init__.py (the Blender add-on entry) ----------------------
pkg = types.ModuleType('ptina')
pkg.__path__ = [os.path.join(REF, 'ptina')]
sys.modules['ptina'] = pkg
with contextlib.redirect_stdout(io.StringIO()):
    import ptina.common as common                      # noqa: E402
    import ptina.things as things                      # noqa: E402
    import ptina.engine.path as enginepath             # noqa: E402
    import ptina.engine.brute as enginebrute           # noqa: E402
    import ptina.sampling as sampling                  # noqa: E402
    import ptina.sampling.sobol as sobolmod            # noqa: E402
    import ptina.tree.lbvh as lbvh                     # noqa: E402
    import ptina.materials.disney as disney            # noqa: E402
    import ptina.materials.microfacet as microfacet    # noqa: E402
    import ptina.tools.matrix                          # noqa: E402  (host NumPy helpers: imported BEFORE the kernel builtins are injected)

for name, mod in list(sys.modules.items()):
    if (name == 'ptina' or name.startswith('ptina.')) and not name.startswith('ptina.tools'):     # tools/* is host NumPy code
        assert getattr(mod, '__file__', None) is None or mod.__file__.startswith(REF), mod.__file__
        mod.__dict__.update(ti.KERNEL_BUILTINS)


def _route_subscript(cls):
    def unpack(idx):
        if isinstance(idx, ti.Matrix):
            return tuple(idx.entries)
        if isinstance(idx, tuple):
            out = []
            for e in idx:
                out.extend(e.entries if isinstance(e, ti.Matrix) else [e])
            return tuple(out)
        return (idx,)

    def getitem(self, idx):
        r = self.subscript(*unpack(idx))
        return r.copy() if isinstance(r, ti.Matrix) else r          # `x = obj[i]` creates a local copy in Taichi

    def setitem(self, idx, val):
        r = self.subscript(*unpack(idx))
        assert isinstance(r, ti._FieldMatrix)
        r._field[r._idx] = val
    cls.__getitem__, cls.__setitem__ = getitem, setitem


from ptina.model import ModelPool          # noqa: E402
from ptina.image import ImagePool, Image   # noqa: E402
from ptina.filmtable import FilmTable      # noqa: E402
from ptina.mtllib import MaterialPool      # noqa: E402
from ptina.light import LightPool          # noqa: E402
from ptina.light.world import WorldLight   # noqa: E402
from ptina.camera import Camera            # noqa: E402
from ptina.tree import BVHTree             # noqa: E402
from ptina.stack import Stack              # noqa: E402
from ptina.geometries import Ray, Box, Face  # noqa: E402
for c in (ModelPool, ImagePool, Image, FilmTable):
    _route_subscript(c)
# the xyz swizzle is read-only in common.py:116-120; `val.xyz /= val.w` (filmtable.py:57) needs the write half Taichi's
# in-place augmented assignment provides
ti.Matrix.xyz = property(lambda v: common.V(v.x, v.y, v.z), lambda v, o: [v._set(i, o.entries[i]) for i in range(3)] and None)


def _model_from_numpy(self, arr, mtlids):          # model.py:53-60 as a NumPy copy
    n = mtlids.shape[0]
    self.nfaces[None] = n
    self.vertices.arr[:n * 24] = np.asarray(arr, np.float32).reshape(-1)
    self.mtlids.arr[:n] = mtlids


def _image_from_numpy(self, id, arr):              # image.py:44-50 as a NumPy copy
    nx, ny, base = int(self.nx[id]), int(self.ny[id]), int(self.base[id])
    self.root.arr[base:base + nx * ny] = np.asarray(arr, np.float32).reshape(nx * ny, 4)


ModelPool.from_numpy = _model_from_numpy
ImagePool.from_numpy = _image_from_numpy
_argsort = np.argsort
lbvh.np = types.SimpleNamespace(**{k: getattr(np, k) for k in ('empty', 'int32', 'ceil')}, log2=lambda x: np.log2(int(x)),   # only feeds a print()
                                argsort=lambda a: _argsort(a, kind='stable'))

from ptina_b200 import scenes  # noqa: E402  (scene GENERATORS only: pure NumPy harness code)


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def fl(x):
    if isinstance(x, ti.Matrix):
        return [float(e) for e in x.entries]
    return float(x)


class RefWorker:
    """The loader calls of exams/benchmark.py:7-26 against the reference singletons."""
    def load_materials(self, m): MaterialPool().load(m)
    def load_images(self, im): quiet(ImagePool().load, im)
    def load_model(self, v, m): ModelPool().load(np.asarray(v), np.asarray(m))
    def build_tree(self): quiet(BVHTree().build)
    def clear_lights(self): LightPool().clear()
    def add_light(self, *a): LightPool().add(*a)
    def set_world_light(self, fac, tex): WorldLight().set(fac, tex)
    def set_camera(self, p): Camera().set_perspective(np.asarray(p))
    def set_size(self, nx, ny): FilmTable().set_size(nx, ny)


def export_tree():
    b = BVHTree()
    n = int(b.n[None])
    return dict(mc=b.mc.arr[:n].copy(), id=b.id.arr[:n].copy(), leaf=b.leaf.arr[:n].copy(), child=b.child.arr[:n - 1].copy(),
                bmin=b.bmin.arr[:n - 1].copy(), bmax=b.bmax.arr[:n - 1].copy())


def trace_pixels(nx, ny):
    """primary ray + closest hit per pixel through the reference's own get_rng / Camera.generate / LinearBVH.intersect"""
    eng = enginepath.PathEngine()
    rays, hits = np.zeros((nx * ny, 6), np.float32), np.zeros((nx * ny, 5), np.float32)
    Stack().set(0)
    for i in range(nx):
        for j in range(ny):
            rng = eng.get_rng(ti.i32(i), ti.i32(j))
            dx, dy = common.random2(rng)
            x = (ti.i32(i) + dx) / FilmTable().nx * 2 - 1
            y = (ti.i32(j) + dy) / FilmTable().ny * 2 - 1
            ray = Camera().generate(x, y)
            rays[i * ny + j] = fl(ray.o) + fl(ray.d)
            ray.d = ray.d.normalized()                       # path.py:28
            h = BVHTree().intersect(ray, -1)
            hits[i * ny + j] = [float(h.hit), float(h.depth), float(h.index), float(h.uv.x), float(h.uv.y)]
    Stack().unset()
    return rays, hits


def render_frames(engine, nframes):
    """engine.render() x nframes exactly as exams/benchmark.py:29-33; returns the Sobol time of each frame and the film"""
    ks = []
    for _ in range(nframes):
        quiet(engine.render)
        ks.append(int(sobolmod.SobolSampler().time[None]))
    nx, ny = int(FilmTable().nx), int(FilmTable().ny)
    film = FilmTable().root.arr[0, :nx * ny].copy().reshape(nx, ny, 4)
    img = FilmTable().get_image()
    out = np.zeros(nx * ny * 3, np.float32)
    FilmTable().fast_export_image(out, 0)
    return np.array(ks), film, np.asarray(img, np.float32), out


def sample_radiance(engine_kind, nx, ny):
    """per-pixel radiance of the CURRENT Sobol point (no film) via the reference's do_render pieces"""
    eng = enginepath.PathEngine() if engine_kind == 'path' else enginebrute.BruteEngine()
    out = np.zeros((nx, ny, 3), np.float32)
    Stack().set(0)
    for i in range(nx):
        for j in range(ny):
            rng = eng.get_rng(ti.i32(i), ti.i32(j))
            dx, dy = common.random2(rng)
            x = (ti.i32(i) + dx) / FilmTable().nx * 2 - 1
            y = (ti.i32(j) + dy) / FilmTable().ny * 2 - 1
            ray = Camera().generate(x, y)
            clr = enginepath.path_trace(ray, rng) if engine_kind == 'path' else eng.trace(ray, rng)
            out[i, j] = fl(clr)
    Stack().unset()
    return out


def scene_arrays(sc):
    """what a parity test needs to rebuild the scene without ptina_b200.scenes (kept so the fixtures stay self-contained)"""
    return dict(vertices=np.asarray(sc['vertices'], np.float64), mtlids=np.asarray(sc['mtlids'], np.int32), pers=np.asarray(sc['pers'], np.float64),
                size=np.asarray(sc['size'], np.int32))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    t00 = time.time()
    quiet(things.init_things)
    print('init_things done', time.time() - t00, flush=True)
    t0 = time.time()
    quiet(enginepath.PathEngine)          # -> SobolSampler(): calc_sobol_vgrid + reset() (64 updates), sobol.py:75-97
    quiet(enginebrute.BruteEngine)
    sob = sobolmod.SobolSampler()
    print(f'SobolSampler ready (time={int(sob.time[None])}) in {time.time() - t0:.1f}s', flush=True)
    G = {}

    # ---- A. Sobol: V grid, P after the 64 reset updates ---------------------------------------------------------------
    np.savez_compressed(os.path.join(HERE, 'sobol.npz'), V=sob.V.arr.copy(), time=np.int32(int(sob.time[None])), P64=sob.P.arr.copy(), X64=sob.X.arr.copy())

    # ---- B/C. hashes, Morton codes, clz -------------------------------------------------------------------------------
    rng = np.random.default_rng(2024)
    ij = rng.integers(0, 2048, (256, 2))
    wh = np.array([int(sampling.wanghash2(ti.i32(int(a)), ti.i32(int(b)))) for a, b in ij], np.int64)
    pts = np.concatenate([rng.random((200, 3)), [[0, 0, 0], [1, 1, 1], [0.5, 0.25, 0.125], [1.5, -0.5, 0.9999999]]]).astype(np.float32)
    mcs = np.array([int(lbvh.morton3D(common.V(ti.f32(float(p[0])), ti.f32(float(p[1])), ti.f32(float(p[2]))))) for p in pts], np.int64)
    xs = np.concatenate([[0, 1, 2, 3, 4, 7, 8, 1 << 29, (1 << 30) - 1], rng.integers(0, 1 << 30, 64)]).astype(np.int64)
    clz = np.array([int(lbvh.clz(ti.i32(int(x)))) for x in xs], np.int64)
    np.savez_compressed(os.path.join(HERE, 'integer.npz'), wh_ij=ij, wh=wh, morton_pts=pts, morton=mcs, clz_x=xs, clz=clz)
    print('integer vectors done', flush=True)

    # ---- G. Disney BSDF eval / sample, microfacet helpers ----------------------------------------------------------------
    m = 400
    params = np.zeros((m, 14), np.float32)
    params[:, 0:3] = rng.uniform(0.05, 1.0, (m, 3)); params[:, 3] = rng.choice([0.0, 0.3, 1.0], m); params[:, 4] = rng.uniform(0.05, 1.0, m)
    params[:, 5] = rng.uniform(0, 1, m); params[:, 6] = rng.uniform(0, 1, m); params[:, 7] = rng.choice([0.0, 0.5], m)
    params[:, 8] = rng.choice([0.0, 0.7], m); params[:, 9] = rng.uniform(0, 1, m); params[:, 10] = rng.choice([0.0, 0.0, 1.0], m)
    params[:, 11] = rng.uniform(0, 1, m); params[:, 12] = rng.choice([0.0, 0.0, 0.5, 1.0], m); params[:, 13] = rng.uniform(1.1, 2.0, m)
    def unit(k):
        v = rng.normal(size=(k, 3)); return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)
    nrm, wi, wo = unit(m), unit(m), unit(m)
    flip = (wi * nrm).sum(1) < 0
    wi[flip] = -wi[flip]
    sign = np.where(rng.random(m) < 0.15, -1.0, 1.0).astype(np.float32)      # the integrator never passes sign < 0, the function handles it
    samp = rng.random((m, 3)).astype(np.float32)
    ev, sm = np.zeros((m, 3), np.float32), np.zeros((m, 7), np.float32)
    F = lambda a: ti.f32(float(a))
    Vv = lambda a: common.V(F(a[0]), F(a[1]), F(a[2]))
    for q in range(m):
        p = params[q]
        mat = disney.Disney(Vv(p[0:3]), *[F(x) for x in p[3:14]])
        ev[q] = fl(mat.brdf(Vv(nrm[q]), F(sign[q]), Vv(wi[q]), Vv(wo[q])))
        s = mat.bounce(Vv(nrm[q]), F(sign[q]), Vv(wi[q]), Vv(samp[q]))
        sm[q] = fl(s.outdir) + [float(s.pdf)] + (fl(s.color) if isinstance(s.color, ti.Matrix) else [float(s.color)] * 3)
    np.savez_compressed(os.path.join(HERE, 'bsdf.npz'), params=params, normal=nrm, sign=sign, wi=wi, wo=wo, samp=samp, eval=ev, sample=sm)
    print('bsdf vectors done', time.time() - t00, flush=True)

    # ---- D/E/F/J. configs 1 and 2: tree, primary rays + hits, frames -------------------------------------------------------------
    for name, size, nfr in (('cornell_boxes', (16, 16), 2), ('cornell_monkey', (12, 12), 2)):
        if args.only and args.only != name:
            continue
        t0 = time.time()
        sc = scenes.CONFIGS[name]()
        sc['size'] = size
        scenes.apply(RefWorker(), sc)
        # a POINT light next to the AREA light so that both kinds and LightPool.hit's first-hit rule are exercised
        from ptina_b200.tools import matrix as mx
        extra = (mx.translate((0.9, 2.6, 0.8)), np.array([4.0, 3.0, 2.0]), 0.25, 'POINT')
        LightPool().add(*extra)
        FilmTable().clear()
        tree = export_tree()
        k_now = int(sob.time[None])
        rays, hits = trace_pixels(*size)                     # at the CURRENT Sobol point (time k_now)
        rad = sample_radiance('path', *size)
        # I. light taps on this light set
        org = rng.uniform([-1.9, 0.1, -1.9], [1.9, 3.9, 1.9], (150, 3)).astype(np.float32)
        dirs = unit(150)
        lh, ls = np.zeros((150, 6), np.float32), np.zeros((150, 8), np.float32)
        lsamp = rng.random((150, 3)).astype(np.float32)
        for q in range(150):
            r = LightPool().hit(Ray(Vv(org[q]), Vv(dirs[q])))
            lh[q] = [float(r.hit), float(r.dis), float(r.pdf)] + fl(r.color)
            s = LightPool().sample(Vv(org[q]), Vv(lsamp[q]))
            ls[q] = [float(s.dis)] + fl(s.dir) + [float(s.pdf)] + fl(s.color)
        ks, film, img, fast = render_frames(enginepath.PathEngine(), nfr)
        np.savez_compressed(os.path.join(HERE, f'{name}.npz'), **scene_arrays(sc), **{'tree_' + k: v for k, v in tree.items()},
                            extra_light_world=extra[0], extra_light_color=extra[1], extra_light_size=np.float32(extra[2]),
                            k_primary=np.int32(k_now), rays=rays, hits=hits, radiance=rad, light_org=org, light_dir=dirs, light_samp=lsamp,
                            light_hit=lh, light_sample=ls, frame_ks=ks, film=film, image=img, fast_export=fast)
        print(f'{name}: golden done in {time.time() - t0:.1f}s (primary k={k_now}, frames {ks.tolist()})', flush=True)

    # ---- K/H. brute engine + textures + environment on the mini matball ----------------------------------------------------------
    if not args.only or args.only == 'mini_matball':
        t0 = time.time()
        sc = scenes.mini_matball()
        scenes.apply(RefWorker(), sc)
        FilmTable().clear()
        tree = export_tree()
        k_now = int(sob.time[None])
        rays, hits = trace_pixels(*sc['size'])
        rad = sample_radiance('brute', *sc['size'])
        mq = 150
        mtl = rng.integers(-1, 5, mq).astype(np.int32)
        uv = rng.uniform(-1.5, 2.5, (mq, 2)).astype(np.float32)
        mg = np.zeros((mq, 14), np.float32)
        for q in range(mq):
            d = MaterialPool().get(ti.i32(int(mtl[q])), common.V(F(uv[q, 0]), F(uv[q, 1])))
            mg[q] = fl(d.basecolor) + [float(getattr(d, s)) for s in ('metallic', 'roughness', 'specular', 'specularTint', 'subsurface', 'sheen', 'sheenTint',
                                                                       'clearcoat', 'clearcoatGloss', 'transmission', 'ior')]
        wd = unit(150)
        wa = np.array([fl(WorldLight().at(Vv(d))) for d in wd], np.float32)
        ks, film, img, fast = render_frames(enginebrute.BruteEngine(), 2)
        np.savez_compressed(os.path.join(HERE, 'mini_matball.npz'), **scene_arrays(sc), **{'tree_' + k: v for k, v in tree.items()},
                            k_primary=np.int32(k_now), rays=rays, hits=hits, radiance=rad, mat_id=mtl, mat_uv=uv, mat_get=mg, world_dir=wd, world_at=wa,
                            frame_ks=ks, film=film, image=img, fast_export=fast)
        print(f'mini_matball: golden done in {time.time() - t0:.1f}s (primary k={k_now}, frames {ks.tolist()})', flush=True)
    print('all golden vectors written in', time.time() - t00, 's')


if __name__ == '__main__':
    main()
