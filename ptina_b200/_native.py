"""ctypes binding of libptina_b200.so (include/ptina_b200.h) and the process-wide context.

The reference keeps its state in metaclass singletons created by `init_things()` (ptina/things.py:12-28,
ptina/common.py:407-413); here they all share ONE native context per process (= one GPU), created
lazily by `context()`.  There is no CPU fallback: if the library cannot be loaded or no CUDA device is
present, every entry point raises.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'libptina_b200.so')

HOST, DEVICE = 0, 1
ENGINE_PATH, ENGINE_BRUTE, ENGINE_PREVIEW, ENGINE_MLT = 0, 1, 2, 3
LIGHT_TYPES = {'POINT': 1, 'AREA': 2}            # light/__init__.py:11
TRAVERSE_AUTO, TRAVERSE_REFERENCE, TRAVERSE_ORDERED, TRAVERSE_ORDERED_EXACT = 0, 1, 2, 3
MODE_PARITY, MODE_FAST = 0, 1

c_f32p = ctypes.POINTER(ctypes.c_float)
c_i32p = ctypes.POINTER(ctypes.c_int32)


class Caps(ctypes.Structure):
    _fields_ = [('max_faces', ctypes.c_int32), ('max_texels', ctypes.c_int32), ('max_materials', ctypes.c_int32),
                ('max_textures', ctypes.c_int32), ('max_lights', ctypes.c_int32), ('max_filmsize', ctypes.c_int32),
                ('max_filmpasses', ctypes.c_int32), ('max_paths', ctypes.c_int64)]


class TreeInfo(ctypes.Structure):
    _fields_ = [('n', ctypes.c_int32), ('aabb_sweeps', ctypes.c_int32), ('valid', ctypes.c_int32), ('depth', ctypes.c_int32),
                ('policy', ctypes.c_int32), ('build_ms', ctypes.c_float), ('list_n', ctypes.c_int32), ('trav_depth', ctypes.c_int32), ('trav_ploc', ctypes.c_int32)]


class Counters(ctypes.Structure):
    _fields_ = [(k, ctypes.c_int64) for k in ('rays', 'extend_rays', 'shadow_rays', 'node_visits', 'box_tests', 'tri_tests',
                                              'paths', 'max_stack')]

    def asdict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


# every symbol include/ptina_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    'ptb_last_error', 'ptb_version', 'ptb_create', 'ptb_destroy', 'ptb_set_stream', 'ptb_set_mode', 'ptb_get_mode', 'ptb_set_option', 'ptb_synchronize', 'ptb_flush',
    'ptb_set_sobol_table', 'ptb_sobol_reset', 'ptb_sobol_get_time', 'ptb_sobol_set_time', 'ptb_sobol_point',
    'ptb_load_model', 'ptb_load_materials', 'ptb_load_images', 'ptb_clear_lights', 'ptb_add_light', 'ptb_set_world_light',
    'ptb_set_camera', 'ptb_build_tree', 'ptb_set_traversal', 'ptb_export_tree', 'ptb_export_traversal', 'ptb_set_size', 'ptb_get_size', 'ptb_clear',
    'ptb_film_ptr', 'ptb_render', 'ptb_render_tile', 'ptb_render_final', 'ptb_render_range', 'ptb_mlt_reset', 'ptb_mlt_set_param', 'ptb_mlt_state', 'ptb_get_image',
    'ptb_fast_export_image', 'ptb_fast_export_gl', 'ptb_get_film', 'ptb_trace_primary', 'ptb_intersect', 'ptb_occluded', 'ptb_eval_bsdf',
    'ptb_sample_bsdf', 'ptb_eval_bsdf_literal', 'ptb_sample_bsdf_literal', 'ptb_material_get', 'ptb_light_hit', 'ptb_light_sample', 'ptb_world_at', 'ptb_normaldist', 'ptb_render_sample',
    'ptb_set_counting', 'ptb_get_counters', 'ptb_reset_counters', 'ptb_get_stage_ms', 'ptb_get_launches', 'ptb_measure_l2', 'ptb_selftest',
]

_lib = None


class NativeError(RuntimeError):
    pass


def library_path():
    return _SO


def load_library():
    """dlopen the in-tree libptina_b200.so.  Raises (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        so = os.environ.get('PTINA_B200_LIB', _SO)        # tuning builds (ptina_b200/build.py --out=...); still no CPU fallback
        if not os.path.exists(so):
            raise NativeError(f'{so} is missing: build it with `python -m ptina_b200.build` '
                              '(ptina_b200 has no CPU fallback)')
        L = ctypes.CDLL(so)
        L.ptb_last_error.restype = ctypes.c_char_p
        for name in SYMBOLS:
            fn = getattr(L, name)
            if name != 'ptb_last_error':
                fn.restype = ctypes.c_int
        _lib = L
    return _lib


def _ptr(a):
    """void* of a NumPy array, a torch tensor (host or CUDA), or None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(ctypes.c_void_p)     # keeps a reference to `a` for the lifetime of the pointer object
    return ctypes.c_void_p(a.data_ptr())   # torch.Tensor


def _is_cuda_tensor(a):
    return (a is not None) and (not isinstance(a, np.ndarray)) and hasattr(a, 'is_cuda') and bool(a.is_cuda)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class _DevView:
    """Zero-copy handle on library-owned device memory (__cuda_array_interface__) so torch can wrap it."""

    def __init__(self, ptr, shape, typestr='<f4'):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2, strides=None)


class Context:
    def __init__(self, device=None, **caps):
        self.L = load_library()
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', '0')) if 'PTB_DEVICE' not in os.environ else int(os.environ['PTB_DEVICE'])
        c = Caps(**caps)
        h = ctypes.c_void_p()
        self.h = None
        self._check(self.L.ptb_create(int(device), ctypes.byref(c), ctypes.byref(h)))
        self.h = h
        self.device = int(device)
        self.caps = c
        self.nfaces = 0
        self.tree = None
        self.use_torch_stream()

    def _check(self, rc):
        if rc != 0:
            raise NativeError(self.L.ptb_last_error().decode())

    def close(self):
        if self.h is not None:
            self.L.ptb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- streams --------------------------------------------------------------------------------
    def use_torch_stream(self):
        """Enqueue on torch's current stream of this device when torch is importable (so torch.cuda.Event timing and
        torch.distributed collectives order correctly against our kernels); otherwise the default stream."""
        try:
            import torch
            if torch.cuda.is_available():
                s = torch.cuda.current_stream(self.device).cuda_stream
                self._check(self.L.ptb_set_stream(self.h, ctypes.c_void_p(s)))
        except ImportError:
            pass

    def synchronize(self):
        self._check(self.L.ptb_synchronize(self.h))

    def set_mode(self, mode):
        """MODE_PARITY (default) or MODE_FAST / 'fast': the shading stage's FMA-contracted, approximate-division build (non-parity)."""
        mode = {'parity': MODE_PARITY, 'fast': MODE_FAST}.get(mode, mode)
        self._check(self.L.ptb_set_mode(self.h, int(mode)))

    def set_option(self, name, value):
        """Scheduling / builder switches that do not change results (see include/ptina_b200.h ptb_set_option)."""
        self._check(self.L.ptb_set_option(self.h, name.encode(), int(value)))

    @property
    def mode(self):
        m = ctypes.c_int()
        self._check(self.L.ptb_get_mode(self.h, ctypes.byref(m)))
        return m.value

    def flush(self):
        """Submit the render() calls recorded so far (they are merged into wavefront batches) without waiting for the device."""
        self._check(self.L.ptb_flush(self.h))

    # ---- sobol -----------------------------------------------------------------------------------
    def set_sobol_table(self, V):
        V = i32(V)
        self.sobol_dim = V.shape[1]
        self._check(self.L.ptb_set_sobol_table(self.h, _ptr(V), V.shape[0], V.shape[1]))

    def sobol_reset(self):
        self._check(self.L.ptb_sobol_reset(self.h))

    @property
    def sobol_time(self):
        t = ctypes.c_int()
        self._check(self.L.ptb_sobol_get_time(self.h, ctypes.byref(t)))
        return t.value

    @sobol_time.setter
    def sobol_time(self, t):
        self._check(self.L.ptb_sobol_set_time(self.h, int(t)))

    def sobol_point(self, k):
        P = np.empty(self.sobol_dim, np.float32)
        self._check(self.L.ptb_sobol_point(self.h, int(k), _ptr(P)))
        return P

    # ---- loaders ---------------------------------------------------------------------------------
    def load_model(self, verts, mtlids):
        """verts [nfaces*3, 8] float32, mtlids [nfaces] int32; NumPy (host) or torch CUDA tensors (device)."""
        dev = _is_cuda_tensor(verts)
        assert dev == _is_cuda_tensor(mtlids), 'vertices and mtlids must live in the same memory space'
        nfaces = int(mtlids.shape[0])
        if dev:      # the library reads raw memory: a float64 or strided tensor would be silently misread
            import torch
            assert verts.dtype == torch.float32 and mtlids.dtype == torch.int32, 'device model data must be float32 / int32'
            assert verts.is_contiguous() and mtlids.is_contiguous(), 'device model data must be contiguous'
            assert verts.device.index == self.device and mtlids.device.index == self.device, 'model tensors live on another GPU'
        else:
            assert verts.dtype == np.float32 and mtlids.dtype == np.int32 and verts.flags.c_contiguous and mtlids.flags.c_contiguous
        assert tuple(verts.shape) == (nfaces * 3, 8), 'vertices must be [nfaces*3, 8]'
        self._check(self.L.ptb_load_model(self.h, _ptr(verts), _ptr(mtlids), nfaces, DEVICE if dev else HOST))
        self._keep = (verts, mtlids)     # keep the sources alive until the async copy is consumed
        self.nfaces = nfaces

    def load_materials(self, fac, tex):
        fac, tex = f32(fac), i32(tex)
        self._check(self.L.ptb_load_materials(self.h, _ptr(fac), _ptr(tex), int(fac.shape[0])))

    def load_images(self, texels, nx, ny, base):
        dev = _is_cuda_tensor(texels)
        nx, ny, base = i32(nx), i32(ny), i32(base)
        n = int(nx.shape[0])
        ntex = int(texels.shape[0]) if n else 0
        self._check(self.L.ptb_load_images(self.h, _ptr(texels) if ntex else None, ctypes.c_int64(ntex), _ptr(nx), _ptr(ny), _ptr(base), n,
                                           DEVICE if dev else HOST))
        self._keep_img = texels

    def clear_lights(self):
        self._check(self.L.ptb_clear_lights(self.h))

    def add_light(self, pos, axes, color, size, type):
        self._check(self.L.ptb_add_light(self.h, _ptr(f32(pos)), _ptr(f32(axes)), _ptr(f32(color)), ctypes.c_float(size), int(type)))

    def set_world_light(self, fac, tex):
        self._check(self.L.ptb_set_world_light(self.h, _ptr(f32(fac)), int(tex)))

    def set_camera(self, v2w, w2v):
        self._check(self.L.ptb_set_camera(self.h, _ptr(f32(v2w)), _ptr(f32(w2v))))

    # ---- tree ------------------------------------------------------------------------------------
    def build_tree(self):
        info = TreeInfo()
        rc = self.L.ptb_build_tree(self.h, ctypes.byref(info))
        self.tree = info
        self._check(rc)
        return info

    def set_traversal(self, policy):
        self._check(self.L.ptb_set_traversal(self.h, int(policy)))

    def export_tree(self):
        n = self.nfaces
        mc, id_, leaf = (np.zeros(n, np.int32) for _ in range(3))
        m = max(n - 1, 0)
        child, bmin, bmax = np.zeros((m, 2), np.int32), np.zeros((m, 3), np.float32), np.zeros((m, 3), np.float32)
        self._check(self.L.ptb_export_tree(self.h, _ptr(mc), _ptr(id_), _ptr(child), _ptr(leaf), _ptr(bmin), _ptr(bmax)))
        return dict(mc=mc, id=id_, child=child, leaf=leaf, bmin=bmin, bmax=bmax)

    def export_traversal(self):
        """The production traversal structure (include/ptina_b200.h ptb_export_traversal)."""
        n = self.nfaces
        nodes = np.zeros((max(n - 1, 0), 16), np.float32)
        lo, hi, gbox = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32), np.zeros((2 * n, 4), np.float32)
        ploc = ctypes.c_int32()
        self._check(self.L.ptb_export_traversal(self.h, _ptr(nodes), _ptr(lo), _ptr(hi), _ptr(gbox), ctypes.byref(ploc)))
        return dict(nodes=nodes, leaf_lo=lo, leaf_hi=hi, gbox=gbox, ploc=bool(ploc.value))

    # ---- film ------------------------------------------------------------------------------------
    def set_size(self, nx, ny):
        self._check(self.L.ptb_set_size(self.h, int(nx), int(ny)))

    def get_size(self):
        nx, ny = ctypes.c_int(), ctypes.c_int()
        self._check(self.L.ptb_get_size(self.h, ctypes.byref(nx), ctypes.byref(ny)))
        return nx.value, ny.value

    def clear(self):
        self._check(self.L.ptb_clear(self.h))

    def film_tensor(self, id=0):
        """torch view [nx*ny, 4] of the running sums of pass `id` (library-owned memory; used for the NCCL reduce)."""
        import torch
        p, n = ctypes.c_void_p(), ctypes.c_int64()
        self._check(self.L.ptb_film_ptr(self.h, int(id), ctypes.byref(p), ctypes.byref(n)))
        return torch.as_tensor(_DevView(p.value, (n.value, 4)), device=f'cuda:{self.device}')

    def _readback(self, fn, id, shape, out=None, device=False):
        if device:
            import torch
            t = out if out is not None else torch.empty(shape, dtype=torch.float32, device=f'cuda:{self.device}')
            self._check(fn(self.h, int(id), _ptr(t), DEVICE))
            return t
        arr = out if out is not None else np.empty(shape, np.float32)
        self._check(fn(self.h, int(id), _ptr(arr), DEVICE if _is_cuda_tensor(arr) else HOST))
        return arr

    def get_image(self, id=0, out=None, device=False):
        nx, ny = self.get_size()
        return self._readback(self.L.ptb_get_image, id, (nx, ny, 4), out, device)

    def get_film(self, id=0, out=None, device=False):
        nx, ny = self.get_size()
        return self._readback(self.L.ptb_get_film, id, (nx, ny, 4), out, device)

    def fast_export_image(self, out, id=0):
        nx, ny = self.get_size()
        assert out.size >= nx * ny * 3 if isinstance(out, np.ndarray) else out.numel() >= nx * ny * 3
        self._check(self.L.ptb_fast_export_image(self.h, int(id), _ptr(out), DEVICE if _is_cuda_tensor(out) else HOST))

    def fast_export_gl(self, gl_buffer, id=0):
        """The same resolve written straight into an OpenGL buffer object of the current GL context (CUDA-GL interop)."""
        self._check(self.L.ptb_fast_export_gl(self.h, int(id), ctypes.c_uint(int(gl_buffer))))

    # ---- render ----------------------------------------------------------------------------------
    def render(self, engine, nsamples=1):
        self._check(self.L.ptb_render(self.h, int(engine), int(nsamples)))

    def render_tile(self, engine, i, j, samples):
        self._check(self.L.ptb_render_tile(self.h, int(engine), int(i), int(j), int(samples)))

    def render_final(self, engine, nsamples):
        self._check(self.L.ptb_render_final(self.h, int(engine), int(nsamples)))

    def render_range(self, engine, k_first, count, stride=1):
        self._check(self.L.ptb_render_range(self.h, int(engine), int(k_first), int(count), int(stride)))

    def render_sample(self, engine, k):
        nx, ny = self.get_size()
        out = np.empty((nx, ny, 3), np.float32)
        self._check(self.L.ptb_render_sample(self.h, int(engine), int(k), _ptr(out)))
        return out

    def mlt_reset(self, seed=0, chain_first=0, chain_count=2**18):
        self._check(self.L.ptb_mlt_reset(self.h, ctypes.c_uint64(seed), int(chain_first), int(chain_count)))

    def mlt_state(self, nchains):
        """nchains must be the chain count of the last mlt_reset (checked by the library)."""
        xn, xo = np.empty((nchains, 32), np.float32), np.empty((nchains, 32), np.float32)
        ln, lo = np.empty((nchains, 3), np.float32), np.empty((nchains, 3), np.float32)
        self._check(self.L.ptb_mlt_state(self.h, int(nchains), _ptr(xn), _ptr(ln), _ptr(xo), _ptr(lo)))
        return dict(X_new=xn, L_new=ln, X_old=xo, L_old=lo)

    def mlt_set_param(self, lsp, sigma):
        self._check(self.L.ptb_mlt_set_param(self.h, ctypes.c_float(lsp), ctypes.c_float(sigma)))

    # ---- taps --------------------------------------------------------------------------------------
    def trace_primary(self, k, trace=True):
        nx, ny = self.get_size()
        npx = nx * ny
        rays = np.empty((npx, 6), np.float32)
        out = dict(rays=rays)
        if trace:
            out.update(hit=np.empty(npx, np.int32), depth=np.empty(npx, np.float32), index=np.empty(npx, np.int32), uv=np.empty((npx, 2), np.float32))
        self._check(self.L.ptb_trace_primary(self.h, int(k), _ptr(rays), _ptr(out.get('hit')), _ptr(out.get('depth')), _ptr(out.get('index')), _ptr(out.get('uv'))))
        return out

    def intersect(self, rays, avoid=None, policy=TRAVERSE_AUTO):
        rays = f32(rays).reshape(-1, 6)
        m = rays.shape[0]
        av = i32(avoid) if avoid is not None else None
        out = dict(hit=np.empty(m, np.int32), depth=np.empty(m, np.float32), index=np.empty(m, np.int32), uv=np.empty((m, 2), np.float32))
        self._check(self.L.ptb_intersect(self.h, _ptr(rays), _ptr(av), m, int(policy), _ptr(out['hit']), _ptr(out['depth']), _ptr(out['index']), _ptr(out['uv'])))
        return out

    def occluded(self, rays, dis, avoid=None, policy=TRAVERSE_AUTO):
        rays, dis = f32(rays).reshape(-1, 6), f32(dis)
        m = rays.shape[0]
        av = i32(avoid) if avoid is not None else None
        out = np.empty(m, np.int32)
        self._check(self.L.ptb_occluded(self.h, _ptr(rays), _ptr(av), _ptr(dis), m, int(policy), _ptr(out)))
        return out

    def _tap(self, fn, a, wa, b, wb, wout, ints=None):
        a = f32(a).reshape(-1, wa)
        m = a.shape[0]
        out = np.empty((m, wout), np.float32)
        if ints is not None:
            self._check(fn(self.h, _ptr(i32(ints)), _ptr(a), m, _ptr(out)))
        elif b is not None:
            self._check(fn(self.h, _ptr(a), _ptr(f32(b).reshape(-1, wb)), m, _ptr(out)))
        else:
            self._check(fn(self.h, _ptr(a), m, _ptr(out)))
        return out

    def eval_bsdf(self, params, geom):
        return self._tap(self.L.ptb_eval_bsdf, params, 14, geom, 10, 3)

    def sample_bsdf(self, params, geom):
        return self._tap(self.L.ptb_sample_bsdf, params, 14, geom, 10, 7)

    def eval_bsdf_literal(self, params, geom):
        return self._tap(self.L.ptb_eval_bsdf_literal, params, 14, geom, 10, 3)

    def sample_bsdf_literal(self, params, geom):
        return self._tap(self.L.ptb_sample_bsdf_literal, params, 14, geom, 10, 7)

    def material_get(self, mtlid, uv):
        return self._tap(self.L.ptb_material_get, uv, 2, None, 0, 14, ints=mtlid)

    def light_hit(self, rays):
        return self._tap(self.L.ptb_light_hit, rays, 6, None, 0, 6)

    def light_sample(self, hitpos_samp):
        return self._tap(self.L.ptb_light_sample, hitpos_samp, 6, None, 0, 8)

    def world_at(self, dirs):
        return self._tap(self.L.ptb_world_at, dirs, 3, None, 0, 3)

    def normaldist(self, samp):
        return self._tap(self.L.ptb_normaldist, samp, 1, None, 0, 1).reshape(-1)

    # ---- counters ------------------------------------------------------------------------------------
    def set_counting(self, count=False, profile=False):
        self._check(self.L.ptb_set_counting(self.h, int(bool(count)) | (int(bool(profile)) << 1)))

    def counters(self):
        c = Counters()
        self._check(self.L.ptb_get_counters(self.h, ctypes.byref(c)))
        return c.asdict()

    def reset_counters(self):
        self._check(self.L.ptb_reset_counters(self.h))

    def stage_ms(self):
        ms = (ctypes.c_float * 5)()
        self._check(self.L.ptb_get_stage_ms(self.h, ms))
        return dict(zip(('raygen', 'extend', 'shade', 'shadow', 'accumulate'), [float(x) for x in ms]))

    def launches(self):
        n = ctypes.c_int64()
        self._check(self.L.ptb_get_launches(self.h, ctypes.byref(n)))
        return n.value

    def selftest(self, what=0, n=1 << 28, seed=1):
        f = ctypes.c_int64()
        self._check(self.L.ptb_selftest(self.h, int(what), ctypes.c_int64(n), ctypes.c_uint64(seed), ctypes.byref(f)))
        return f.value

    def measure_l2(self, mbytes=64, iters=20):
        g = ctypes.c_float()
        self._check(self.L.ptb_measure_l2(self.h, int(mbytes), int(iters), ctypes.byref(g)))
        return g.value


_ctx = None


def context(create=True, **caps):
    """The process-wide context (the reference's singleton state).  Created on first use."""
    global _ctx
    if _ctx is None and create:
        _ctx = Context(**caps)
    if _ctx is None:
        raise NativeError('ptina_b200 is not initialised: call init_things() / worker.init() first')
    return _ctx


def shutdown():
    global _ctx
    if _ctx is not None:
        _ctx.close()
        _ctx = None
