"""TEST INFRASTRUCTURE -- the CPU oracle.  Not product code.

ctypes front-end of oracle/ptina_oracle.cpp plus literal restatements of the host-side halves of the
reference's loaders (the parts PTina runs in Python/NumPy before any kernel):

  MaterialPool.load / ParameterPair.load   ptina/mtllib.py:15-28, 58-77
  ImagePool.load / load_one, allocators    ptina/image.py:69-95, ptina/allocator.py:6-53
  LightPool.add / clear                    ptina/light/__init__.py:31-49
  Camera.set_perspective                   ptina/camera.py:19-22
  ModelPool.load                           ptina/model.py:62-86

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'libptina_oracle.so')
_lib = None

ENGINE_PATH, ENGINE_BRUTE = 0, 1


def build(force=False):
    src = os.path.join(_HERE, 'ptina_oracle.cpp')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE], env=dict(os.environ, CXX='', CC=''))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.ora_create.restype = ctypes.c_void_p
        for name in ('ora_destroy', 'ora_set_sobol', 'ora_sobol_point', 'ora_load_model', 'ora_load_materials',
                     'ora_load_images', 'ora_clear_lights', 'ora_add_light', 'ora_set_world_light', 'ora_set_camera',
                     'ora_set_size', 'ora_clear', 'ora_build_tree', 'ora_export_tree', 'ora_validate_tree',
                     'ora_intersect', 'ora_primary', 'ora_render', 'ora_render_sample', 'ora_trace_from_samples',
                     'ora_get_film', 'ora_get_image', 'ora_fast_export_image', 'ora_eval_bsdf', 'ora_sample_bsdf',
                     'ora_material_get', 'ora_light_hit', 'ora_light_sample', 'ora_world_at'):
            getattr(L, name).restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Oracle:
    """One reference 'process': the singletons of ptina/things.py:12-28 rolled into one object."""

    def __init__(self, sobol=True):
        self.L = lib()
        self.h = ctypes.c_void_p(self.L.ora_create())
        self.nx = self.ny = 0
        self.nfaces = 0
        self.sobol_time = 0
        if sobol:
            from .sobol_table import vgrid_i32
            V = vgrid_i32()
            self.dim = V.shape[1]
            assert self.L.ora_set_sobol(self.h, _p(V), V.shape[0], V.shape[1]) == 0
            self.sobol_time = 64      # SobolSampler.reset(): skip=64 updates (sobol.py:92-97)
        # LightPool default light (light/__init__.py:22-28)
        self.clear_lights()
        self.add_light(np.array([[1, 0, 0, 1], [0, 1, 0, 2], [0, 0, 1, 3], [0, 0, 0, 1.0]]), np.array([32., 32, 32]), 0.5, 'POINT')
        self._zero_axes_default = True
        self.set_camera(np.eye(4))

    def __del__(self):
        try:
            self.L.ora_destroy(self.h)
        except Exception:
            pass

    # ---- ModelPool.load (model.py:62-86) -------------------------------------------------------
    def load_model(self, arr, mtlids=None):
        arr = np.asarray(arr)
        if arr.dtype == np.float64:
            arr = arr.astype(np.float32)
        assert arr.shape[0] % 3 == 0
        if mtlids is None:
            mtlids = -np.ones(arr.shape[0] // 3, dtype=np.int32)
        else:
            assert mtlids.shape[0] == arr.shape[0] // 3
        assert mtlids.shape[0] < 2**21, 'too many faces'
        a, m = _f32(arr), _i32(mtlids)
        self.nfaces = m.shape[0]
        self.L.ora_load_model(self.h, _p(a), _p(m), self.nfaces)

    # ---- MaterialPool.load (mtllib.py:58-77) + ParameterPair.load (mtllib.py:15-28) ---------------
    def load_materials(self, materials):
        fac = np.zeros((64, 12, 4), np.float32)   # zero-initialised Taichi fields
        tex = np.zeros((64, 12), np.int32)
        for i, material in enumerate(materials):
            for k, (f, t) in zip(range(12), material):
                if f is None:
                    f = 1.0
                if isinstance(f, np.ndarray):
                    f = list(f) if len(f.shape) else float(f)
                if not isinstance(f, (tuple, list)):
                    f = [f, f, f, f]
                if isinstance(f, (tuple, list)) and len(f) == 3:
                    f = list(f) + [1.0]
                fac[i, k] = f
                tex[i, k] = t
        self.L.ora_load_materials(self.h, _p(fac), _p(tex), 64)

    # ---- ImagePool.load (image.py:69-95), MemoryAllocator first-fit (allocator.py:6-26) -----------
    def load_images(self, images):
        free = [(0, 2**22)]
        nx_, ny_, base_, chunks = [], [], [], []
        for arr in images:
            arr = np.asarray(arr)
            if arr.dtype == np.uint8:
                arr = arr.astype(np.float32) / 255
            nx, ny = arr.shape[0], arr.shape[1]
            if len(arr.shape) == 2:
                arr = arr[:, :, None]
            if arr.shape[2] == 1:
                arr = np.stack([arr[:, :, 0]] * 3, axis=2)
            if arr.shape[2] == 3:
                arr = np.concatenate([arr, np.ones((nx, ny, 1))], axis=2)
            if len(nx_) >= 64:
                raise RuntimeError('Out of ID!')
            for i, (cb, cs) in enumerate(free):
                if cs >= nx * ny:
                    del free[i]
                    if cs != nx * ny:
                        free.insert(i, (cb + nx * ny, cs - nx * ny))
                    base = cb
                    break
            else:
                raise RuntimeError('Out of memory!')
            nx_.append(nx); ny_.append(ny); base_.append(base); chunks.append(_f32(arr).reshape(nx * ny, 4))
        ntex = (base_[-1] + nx_[-1] * ny_[-1]) if images else 0
        tex = np.zeros((max(ntex, 1), 4), np.float32)
        for b, c in zip(base_, chunks):
            tex[b:b + c.shape[0]] = c
        self.L.ora_load_images(self.h, _p(tex), ctypes.c_longlong(ntex), _p(_i32(nx_ or [0])), _p(_i32(ny_ or [0])),
                               _p(_i32(base_ or [0])), len(nx_))

    # ---- LightPool (light/__init__.py:31-49) ------------------------------------------------------
    def clear_lights(self):
        self.L.ora_clear_lights(self.h)

    def add_light(self, world, color, size, type):
        world = np.asarray(world, dtype=np.float64)
        pos = world @ np.array([0, 0, 0, 1])
        pos = pos[:3] / pos[3]
        axes = world[:3, :3]
        return self.L.ora_add_light(self.h, _p(_f32(pos)), _p(_f32(axes)), _p(_f32(np.asarray(color))),
                                    ctypes.c_float(size), {'POINT': 1, 'AREA': 2}[type])

    def set_world_light(self, fac, tex):
        self.L.ora_set_world_light(self.h, _p(_f32(fac)), int(tex))

    # ---- Camera.set_perspective (camera.py:19-22) ---------------------------------------------------
    def set_camera(self, pers):
        invpers = np.linalg.inv(np.asarray(pers, dtype=np.float64))
        self.L.ora_set_camera(self.h, _p(_f32(invpers)))

    def set_size(self, nx, ny):
        self.nx, self.ny = nx, ny
        self.L.ora_set_size(self.h, nx, ny)

    def clear(self):
        self.L.ora_clear(self.h)

    # ---- BVHTree().build() (lbvh.py:297-305) ----------------------------------------------------------
    def build_tree(self):
        r = self.L.ora_build_tree(self.h)
        if r < 0:
            raise RuntimeError('AABB step never stop! hierarchy corrupted?')
        return r

    def export_tree(self):
        n = self.nfaces
        mc, id_, leaf = (np.zeros(n, np.int32) for _ in range(3))
        child = np.zeros((max(n - 1, 1), 2), np.int32)
        bmin = np.zeros((max(n - 1, 1), 3), np.float32)
        bmax = np.zeros((max(n - 1, 1), 3), np.float32)
        self.L.ora_export_tree(self.h, _p(mc), _p(id_), _p(child), _p(leaf), _p(bmin), _p(bmax))
        return dict(mc=mc, id=id_, child=child[:max(n - 1, 0)], leaf=leaf, bmin=bmin[:max(n - 1, 0)], bmax=bmax[:max(n - 1, 0)])

    def validate_tree(self):
        return self.L.ora_validate_tree(self.h)

    # ---- taps ---------------------------------------------------------------------------------------
    def sobol_point(self, k):
        P = np.zeros(self.dim, np.float32)
        self.L.ora_sobol_point(self.h, int(k), _p(P))
        return P

    def intersect(self, rays, avoid=None, policy=0, counters=False):
        rays = _f32(rays).reshape(-1, 6)
        m = rays.shape[0]
        av = _i32(avoid) if avoid is not None else None
        hit, index = np.zeros(m, np.int32), np.zeros(m, np.int32)
        depth, uv = np.zeros(m, np.float32), np.zeros((m, 2), np.float32)
        cnt = np.zeros(5, np.int64)
        self.L.ora_intersect(self.h, _p(rays), _p(av), m, policy, _p(hit), _p(depth), _p(index), _p(uv), _p(cnt))
        out = dict(hit=hit, depth=depth, index=index, uv=uv)
        if counters:
            out['counters'] = dict(zip(('rays', 'node_visits', 'box_tests', 'tri_tests', 'max_stack'), cnt.tolist()))
        return out

    def primary(self, k, trace=True):
        npx = self.nx * self.ny
        rays = np.zeros((npx, 6), np.float32)
        hit, index = np.zeros(npx, np.int32), np.zeros(npx, np.int32)
        depth, uv = np.zeros(npx, np.float32), np.zeros((npx, 2), np.float32)
        self.L.ora_primary(self.h, int(k), _p(rays), _p(hit) if trace else None, _p(depth), _p(index), _p(uv))
        return dict(rays=rays, hit=hit, depth=depth, index=index, uv=uv)

    def render_sample(self, engine, k):
        out = np.zeros((self.nx * self.ny, 3), np.float32)
        self.L.ora_render_sample(self.h, engine, int(k), _p(out))
        return out.reshape(self.nx, self.ny, 3)

    def render(self, engine=ENGINE_PATH, nsamples=1, k_first=None, nthreads=0):
        """`nsamples` calls of Engine.render(): Sobol point index advances by one per call (sobol.py:99-105)."""
        if k_first is None:
            k_first = self.sobol_time + 1
            self.sobol_time += nsamples
        cnt = np.zeros(5, np.int64)
        r = self.L.ora_render(self.h, engine, int(k_first), int(nsamples), int(nthreads), _p(cnt))
        assert r == 0
        return dict(zip(('rays', 'node_visits', 'box_tests', 'tri_tests', 'max_stack'), cnt.tolist()))

    def render_preview(self, nsamples=1, k_first=None):
        """`nsamples` calls of PreviewEngine.render() (preview.py:19-41): albedo -> film 1, normal -> film 2."""
        if k_first is None:
            k_first = self.sobol_time + 1
            self.sobol_time += nsamples
        assert self.L.ora_render_preview(self.h, int(k_first), int(nsamples)) == 0

    def render_tile(self, engine, i, j, samples):
        """PathEngine.render_tile (path.py:96-118): one Sobol update, then samples 0..min(samples, 63) of the 64x64 tile (i, j)."""
        self.sobol_time += 1
        assert self.L.ora_render_tile(self.h, engine, int(self.sobol_time), int(i), int(j), int(samples)) == 0

    def render_final(self, engine, nsamples):
        """path.py:120-128."""
        for i in range((self.nx + 63) // 64):
            for j in range((self.ny + 63) // 64):
                samples = nsamples
                while samples > 0:
                    self.render_tile(engine, i, j, samples)
                    samples -= 64

    def trace_from_samples(self, X):
        X = _f32(X)
        out = np.zeros((X.shape[0], 3), np.float32)
        self.L.ora_trace_from_samples(self.h, _p(X), X.shape[0], X.shape[1], _p(out))
        return out

    def get_film(self, id=0):
        out = np.zeros((self.nx, self.ny, 4), np.float32)
        self.L.ora_get_film(self.h, id, _p(out))
        return out

    def get_image(self, id=0):
        out = np.zeros((self.nx, self.ny, 4), np.float32)
        self.L.ora_get_image(self.h, id, _p(out))
        return out

    def fast_export_image(self, out, id=0):
        assert out.dtype == np.float32 and out.size >= self.nx * self.ny * 3
        self.L.ora_fast_export_image(self.h, id, _p(out))

    def material_get(self, mtlid, uv):
        mtlid, uv = _i32(mtlid), _f32(uv).reshape(-1, 2)
        out = np.zeros((mtlid.shape[0], 14), np.float32)
        self.L.ora_material_get(self.h, _p(mtlid), _p(uv), mtlid.shape[0], _p(out))
        return out

    def light_hit(self, rays):
        rays = _f32(rays).reshape(-1, 6)
        out = np.zeros((rays.shape[0], 6), np.float32)
        self.L.ora_light_hit(self.h, _p(rays), rays.shape[0], _p(out))
        return out

    def light_sample(self, hitpos_samp):
        a = _f32(hitpos_samp).reshape(-1, 6)
        out = np.zeros((a.shape[0], 8), np.float32)
        self.L.ora_light_sample(self.h, _p(a), a.shape[0], _p(out))
        return out

    def world_at(self, dirs):
        a = _f32(dirs).reshape(-1, 3)
        out = np.zeros((a.shape[0], 3), np.float32)
        self.L.ora_world_at(self.h, _p(a), a.shape[0], _p(out))
        return out


def eval_bsdf(params, geom):
    params, geom = _f32(params).reshape(-1, 14), _f32(geom).reshape(-1, 10)
    out = np.zeros((params.shape[0], 3), np.float32)
    lib().ora_eval_bsdf(_p(params), _p(geom), params.shape[0], _p(out))
    return out


def sample_bsdf(params, geom):
    params, geom = _f32(params).reshape(-1, 14), _f32(geom).reshape(-1, 10)
    out = np.zeros((params.shape[0], 7), np.float32)
    lib().ora_sample_bsdf(_p(params), _p(geom), params.shape[0], _p(out))
    return out


def normaldist(samp):
    samp = _f32(samp).reshape(-1)
    out = np.zeros_like(samp)
    lib().ora_normaldist(_p(samp), samp.shape[0], _p(out))
    return out


def erfinv(x):
    x = _f32(x).reshape(-1)
    out = np.zeros_like(x)
    lib().ora_erfinv(_p(x), x.shape[0], _p(out))
    return out


def wanghash2(x, y):
    return lib().ora_wanghash2(int(x), int(y))


def morton3d(x, y, z):
    return lib().ora_morton3d(ctypes.c_float(x), ctypes.c_float(y), ctypes.c_float(z))


def clz(x):
    return lib().ora_clz(int(x))


def num_threads():
    return lib().ora_num_threads()
