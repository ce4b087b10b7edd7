"""TEST INFRASTRUCTURE -- a tiny interpreter-style stand-in for the `taichi` module (Taichi ~0.7 API surface used by
archibate/ptina), so that the reference's OWN Python sources can be imported from /root/reference and executed here,
in plain CPython, to produce golden vectors (tests/golden/make_golden.py).  Taichi itself is not installable in this
image.

It is NOT a Taichi re-implementation: kernels and funcs simply run as Python, with scalar/vector/field types that carry
Taichi's default-precision semantics:
  * float = IEEE binary32 (every operation rounds to f32; literals are rounded to f32 before use), int = wrapping i32,
    u32 for ti.cast(x, ti.u32); int/int true division -> f32; `//`, `%` floor-style; `int()` truncates (NaN -> INT_MIN as
    x86 cvttss2si); min/max = fminf/fmaxf; x ** n for small integer n = exponentiation by squaring (alg_simp).
  * Matrix @ accumulates k ascending; dot = left-to-right sum; normalized() = (1 / (norm + 0)) * v; cross as written in
    Taichi's matrix.py.
  * transcendental functions call glibc's libm (sinf, cosf, atan2f, logf, powf, expf) -- what Taichi's CPU backend calls.
  * no fast-math rewrites (Taichi's default fast_math=True may turn x / const into x * (1/const) and contract FMAs; the
    canonical semantics of this project is strict IEEE in source order).
  * ti.func copies Matrix arguments (Taichi passes by value); fields are NumPy arrays; parallel fors run sequentially.
"""
import builtins
import ctypes
import itertools
import math
import struct
import types

import numpy as np

_tinahacked = 1          # makes ptina/common.py skip its block of Taichi monkey-patches (they target real Taichi internals)
pi = math.pi
tau = math.tau

_libm = ctypes.CDLL('libm.so.6')
for _n in ('sinf', 'cosf', 'logf', 'expf', 'tanf', 'asinf', 'acosf'):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float]
for _n in ('atan2f', 'powf'):
    getattr(_libm, _n).restype = ctypes.c_float
    getattr(_libm, _n).argtypes = [ctypes.c_float, ctypes.c_float]

_pack = struct.Struct('f')


def _rf(x):
    """round a Python float (double) to the nearest binary32, returned as a Python float"""
    try:
        return _pack.unpack(_pack.pack(x))[0]
    except OverflowError:
        return math.copysign(math.inf, x)


def _wi(x):
    x &= 0xFFFFFFFF
    return x - 0x100000000 if x & 0x80000000 else x


class _Scalar:
    __slots__ = ('v',)

    def __repr__(self):
        return f'{type(self).__name__}({self.v!r})'

    def __hash__(self):
        return hash(self.v)

    def __bool__(self):
        return self.v != 0


def _is_float_like(o):
    return isinstance(o, (f32, builtins.float, np.floating))


def _pyval(o):
    if isinstance(o, _Scalar):
        return o.v
    if isinstance(o, (np.integer,)):
        return builtins.int(o)
    if isinstance(o, (np.floating,)):
        return builtins.float(o)
    return o


class f32(_Scalar):
    __slots__ = ()

    def __init__(self, v=0.0):
        self.v = _rf(builtins.float(_pyval(v)))

    @staticmethod
    def _o(o):
        return _rf(builtins.float(_pyval(o)))

    def __add__(self, o): return _f(self.v + f32._o(o)) if _num(o) else NotImplemented
    def __radd__(self, o): return _f(f32._o(o) + self.v) if _num(o) else NotImplemented
    def __sub__(self, o): return _f(self.v - f32._o(o)) if _num(o) else NotImplemented
    def __rsub__(self, o): return _f(f32._o(o) - self.v) if _num(o) else NotImplemented
    def __mul__(self, o): return _f(self.v * f32._o(o)) if _num(o) else NotImplemented
    def __rmul__(self, o): return _f(f32._o(o) * self.v) if _num(o) else NotImplemented
    def __truediv__(self, o): return _f(_div(self.v, f32._o(o))) if _num(o) else NotImplemented
    def __rtruediv__(self, o): return _f(_div(f32._o(o), self.v)) if _num(o) else NotImplemented
    def __floordiv__(self, o): return _f(math.floor(_div(self.v, f32._o(o)))) if _num(o) else NotImplemented
    def __mod__(self, o):
        if not _num(o): return NotImplemented
        b = f32._o(o)
        r = math.fmod(self.v, b)
        if r != 0 and (r < 0) != (b < 0):
            r += b
        return _f(r)
    def __neg__(self): return _f(-self.v)
    def __pos__(self): return self
    def __abs__(self): return _f(abs(self.v))
    def __pow__(self, o):
        e = _pyval(o)
        if isinstance(e, builtins.int) and not isinstance(e, bool) and 0 < e <= 32:
            return _ipow(self, e)
        return _f(_libm.powf(self.v, builtins.float(e)))
    def __rpow__(self, o): return _f(_libm.powf(f32._o(o), self.v))
    def __lt__(self, o): return self.v < f32._o(o)
    def __le__(self, o): return self.v <= f32._o(o)
    def __gt__(self, o): return self.v > f32._o(o)
    def __ge__(self, o): return self.v >= f32._o(o)
    def __eq__(self, o): return _num(o) and self.v == f32._o(o)
    def __ne__(self, o): return (not _num(o)) or self.v != f32._o(o)
    def __float__(self): return self.v
    def __int__(self): return _trunc_i32(self.v)
    __hash__ = _Scalar.__hash__


def _div(a, b):
    if b == 0:
        if a == 0 or a != a:
            return math.nan
        return math.copysign(math.inf, a) * math.copysign(1.0, b)
    return a / b


def _f(x):
    r = f32.__new__(f32)
    r.v = _rf(x)
    return r


def _ipow(a, n):
    # Taichi alg_simp exponent_n_optimize: exponentiation by squaring
    result, p, cur = None, a, 1
    while True:
        if n & cur:
            result = p if result is None else result * p
        cur <<= 1
        if cur > n:
            break
        p = p * p
    return result


def _trunc_i32(x):
    if x != x or x >= 2147483648.0 or x < -2147483648.0:
        return -2147483648
    return builtins.int(x)


class _Int(_Scalar):
    __slots__ = ()
    signed = True

    @classmethod
    def _mk(cls, x):
        r = cls.__new__(cls)
        r.v = _wi(x) if cls.signed else (x & 0xFFFFFFFF)
        return r

    def __init__(self, v=0):
        v = _pyval(v)
        if isinstance(v, builtins.float):
            v = _trunc_i32(v)
        self.v = _wi(builtins.int(v)) if self.signed else (builtins.int(v) & 0xFFFFFFFF)

    def _res(self, o):
        # C-like promotion: u32 wins over i32
        return u32 if (isinstance(self, u32) or isinstance(o, u32)) else i32

    def _bin(self, o, fn, rev=False):
        if _is_float_like(o):
            a, b = builtins.float(self.v), f32._o(o)
            return None
        return None

    def __add__(self, o):
        if _is_float_like(o): return _f(_rf(builtins.float(self.v)) + f32._o(o))
        return self._res(o)._mk(self.v + _ival(o)) if _num(o) else NotImplemented
    __radd__ = __add__
    def __sub__(self, o):
        if _is_float_like(o): return _f(_rf(builtins.float(self.v)) - f32._o(o))
        return self._res(o)._mk(self.v - _ival(o)) if _num(o) else NotImplemented
    def __rsub__(self, o):
        if _is_float_like(o): return _f(f32._o(o) - _rf(builtins.float(self.v)))
        return self._res(o)._mk(_ival(o) - self.v) if _num(o) else NotImplemented
    def __mul__(self, o):
        if _is_float_like(o): return _f(_rf(builtins.float(self.v)) * f32._o(o))
        return self._res(o)._mk(self.v * _ival(o)) if _num(o) else NotImplemented
    __rmul__ = __mul__
    def __truediv__(self, o):
        return _f(_div(_rf(builtins.float(self.v)), f32._o(o))) if _num(o) else NotImplemented
    def __rtruediv__(self, o):
        return _f(_div(f32._o(o), _rf(builtins.float(self.v)))) if _num(o) else NotImplemented
    def __floordiv__(self, o):
        if _is_float_like(o): return _f(math.floor(_div(builtins.float(self.v), f32._o(o))))
        return self._res(o)._mk(self.v // _ival(o))
    def __rfloordiv__(self, o):
        return self._res(o)._mk(_ival(o) // self.v)
    def __mod__(self, o):
        if _is_float_like(o): return f32(self.v) % o
        return self._res(o)._mk(self.v % _ival(o))
    def __rmod__(self, o):
        return self._res(o)._mk(_ival(o) % self.v)
    def __and__(self, o): return self._res(o)._mk((self.v & 0xFFFFFFFF) & (_ival(o) & 0xFFFFFFFF))
    __rand__ = __and__
    def __or__(self, o): return self._res(o)._mk((self.v & 0xFFFFFFFF) | (_ival(o) & 0xFFFFFFFF))
    __ror__ = __or__
    def __xor__(self, o): return self._res(o)._mk((self.v & 0xFFFFFFFF) ^ (_ival(o) & 0xFFFFFFFF))
    __rxor__ = __xor__
    def __lshift__(self, o): return type(self)._mk((self.v & 0xFFFFFFFF) << (_ival(o) & 31))
    def __rshift__(self, o): return type(self)._mk(self.v >> (_ival(o) & 31))     # i32: arithmetic, u32: logical (v >= 0)
    def __rlshift__(self, o): return i32._mk(_ival(o) << (self.v & 31))
    def __rrshift__(self, o): return i32._mk(_ival(o) >> (self.v & 31))
    def __invert__(self): return type(self)._mk(~self.v)
    def __neg__(self): return type(self)._mk(-self.v)
    def __pos__(self): return self
    def __abs__(self): return type(self)._mk(abs(self.v))
    def __pow__(self, o):
        e = _pyval(o)
        if isinstance(e, builtins.int) and 0 < e <= 32:
            return _ipow(self, e)
        return _f(_libm.powf(builtins.float(self.v), builtins.float(e)))
    def __lt__(self, o): return self.v < _cmpval(o)
    def __le__(self, o): return self.v <= _cmpval(o)
    def __gt__(self, o): return self.v > _cmpval(o)
    def __ge__(self, o): return self.v >= _cmpval(o)
    def __eq__(self, o): return _num(o) and self.v == _cmpval(o)
    def __ne__(self, o): return (not _num(o)) or self.v != _cmpval(o)
    def __index__(self): return self.v
    def __int__(self): return self.v
    def __float__(self): return builtins.float(self.v)
    __hash__ = _Scalar.__hash__


class i32(_Int):
    __slots__ = ()
    signed = True


class u32(_Int):
    __slots__ = ()
    signed = False


def _ival(o):
    o = _pyval(o)
    if isinstance(o, builtins.float):
        raise TypeError('float used where an integer is required')
    return builtins.int(o)


def _cmpval(o):
    o = _pyval(o)
    return _rf(o) if isinstance(o, builtins.float) else o


def _num(o):
    return isinstance(o, (_Scalar, builtins.int, builtins.float, np.integer, np.floating)) and not isinstance(o, Matrix)


f64 = f32          # never used with real f64 data on the hot path
i64 = i32
u8 = u16 = u64 = u32
i8 = i16 = i32


# ------------------------------------------------------------------------------------------------------------------
# Matrix / Vector
# ------------------------------------------------------------------------------------------------------------------
class Matrix:
    is_taichi_class = True

    def __init__(self, rows=None, n=None, m=None, _entries=None):
        if _entries is not None:
            self.entries, self.n, self.m = list(_entries), n, m
            return
        rows = list(rows)
        if rows and isinstance(rows[0], (list, tuple)):
            self.n, self.m = len(rows), len(rows[0])
            self.entries = [x for r in rows for x in r]
        elif rows and isinstance(rows[0], Matrix):            # list of column vectors is handled by cols(); rows of vectors:
            self.n, self.m = len(rows), rows[0].n
            self.entries = [x for r in rows for x in r.entries]
        else:
            self.n, self.m = len(rows), 1
            self.entries = rows

    # -- construction helpers
    @staticmethod
    def cols(cols):
        n, m = cols[0].n, len(cols)
        return Matrix(_entries=[cols[j].entries[i] for i in range(n) for j in range(m)], n=n, m=m)

    @staticmethod
    def rows(rows):
        return Matrix([r.entries for r in rows])

    @staticmethod
    def empty(n, m):
        return Matrix(_entries=[None] * (n * m), n=n, m=m)

    @staticmethod
    def zero(dt, n, m=1):
        return Matrix(_entries=[dt(0)] * (n * m), n=n, m=m)

    @staticmethod
    def unit(n, i, dt=None):
        return Matrix(_entries=[1 if j == i else 0 for j in range(n)], n=n, m=1)

    def copy(self):
        return Matrix(_entries=self.entries, n=self.n, m=self.m)

    # -- element access
    def _lin(self, idx):
        if isinstance(idx, tuple):
            i, j = idx if len(idx) == 2 else (idx[0], 0)
            return builtins.int(_pyval(i)) * self.m + builtins.int(_pyval(j))
        return builtins.int(_pyval(idx)) * self.m if self.m == 1 else builtins.int(_pyval(idx))

    def __getitem__(self, idx):
        return self.entries[self._lin(idx)]

    def __setitem__(self, idx, val):
        self._set(self._lin(idx), val)

    def _set(self, k, val):
        self.entries[k] = val

    def __call__(self, i, j=0):
        return self.entries[i * self.m + j]

    def __iter__(self):
        return iter(self.entries)

    def __len__(self):
        return len(self.entries)

    x = property(lambda s: s.entries[0], lambda s, v: s._set(0, v))
    y = property(lambda s: s.entries[1], lambda s, v: s._set(1, v))
    z = property(lambda s: s.entries[2], lambda s, v: s._set(2, v))
    w = property(lambda s: s.entries[3], lambda s, v: s._set(3, v))

    # -- element-wise arithmetic
    def _ew(self, o, fn):
        if isinstance(o, Matrix):
            assert len(o.entries) == len(self.entries), 'shape mismatch'
            return Matrix(_entries=[fn(_as_scalar(a), b) for a, b in zip(self.entries, o.entries)], n=self.n, m=self.m)
        return Matrix(_entries=[fn(_as_scalar(a), o) for a in self.entries], n=self.n, m=self.m)

    def __add__(self, o): return self._ew(o, lambda a, b: a + b)
    def __radd__(self, o): return self._ew(o, lambda a, b: b + a)
    def __sub__(self, o): return self._ew(o, lambda a, b: a - b)
    def __rsub__(self, o): return self._ew(o, lambda a, b: b - a)
    def __mul__(self, o): return self._ew(o, lambda a, b: a * b)
    def __rmul__(self, o): return self._ew(o, lambda a, b: b * a)
    def __truediv__(self, o): return self._ew(o, lambda a, b: _as_scalar(a) / b)
    def __rtruediv__(self, o): return self._ew(o, lambda a, b: _as_scalar(b) / a)
    def __floordiv__(self, o): return self._ew(o, lambda a, b: _as_scalar(a) // b)
    def __mod__(self, o): return self._ew(o, lambda a, b: _as_scalar(a) % b)
    def __pow__(self, o): return self._ew(o, lambda a, b: _as_scalar(a) ** b)
    def __and__(self, o): return self._ew(o, lambda a, b: _as_scalar(a) & b)
    def __or__(self, o): return self._ew(o, lambda a, b: _as_scalar(a) | b)
    def __xor__(self, o): return self._ew(o, lambda a, b: _as_scalar(a) ^ b)
    def __lshift__(self, o): return self._ew(o, lambda a, b: _as_scalar(a) << b)
    def __rshift__(self, o): return self._ew(o, lambda a, b: _as_scalar(a) >> b)
    def __neg__(self): return Matrix(_entries=[-a for a in self.entries], n=self.n, m=self.m)
    def __pos__(self): return self
    def __abs__(self): return Matrix(_entries=[abs(a) for a in self.entries], n=self.n, m=self.m)
    def __lt__(self, o): return self._ew(o, lambda a, b: a < b)
    def __le__(self, o): return self._ew(o, lambda a, b: a <= b)
    def __gt__(self, o): return self._ew(o, lambda a, b: a > b)
    def __ge__(self, o): return self._ew(o, lambda a, b: a >= b)
    def __eq__(self, o): return self._ew(o, lambda a, b: a == b)
    def __ne__(self, o): return self._ew(o, lambda a, b: a != b)
    __hash__ = None

    def __matmul__(self, o):
        assert self.m == o.n
        out = []
        for i in range(self.n):
            for j in range(o.m):
                acc = self(i, 0) * o(0, j)
                for k in range(1, o.n):
                    acc = acc + self(i, k) * o(k, j)
                out.append(acc)
        return Matrix(_entries=out, n=self.n, m=o.m)

    # -- reductions
    def sum(self):
        r = self.entries[0]
        for e in self.entries[1:]:
            r = r + e
        return r

    def dot(self, o): return (self * o).sum()
    def norm_sqr(self): return (self ** 2).sum()
    def norm(self, eps=0): return sqrt(self.norm_sqr() + eps)
    def normalized(self, eps=0):
        invlen = 1 / (self.norm() + eps)
        return invlen * self
    def cross(self, o):
        a, b = self, o
        return Matrix(_entries=[a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x], n=3, m=1)
    def any(self): return builtins.any(builtins.bool(e != 0) for e in self.entries)
    def all(self): return builtins.all(builtins.bool(e != 0) for e in self.entries)
    def max(self): return ti_max(*self.entries)
    def min(self): return ti_min(*self.entries)
    def transpose(self): return Matrix(_entries=[self(i, j) for j in range(self.m) for i in range(self.n)], n=self.m, m=self.n)
    def cast(self, dt): return Matrix(_entries=[cast(e, dt) for e in self.entries], n=self.n, m=self.m)
    def to_numpy(self):
        return np.array([builtins.float(e) for e in self.entries], dtype=np.float32).reshape((self.n, self.m) if self.m > 1 else (self.n,))
    def variable(self): return self.copy()

    def __repr__(self):
        return f'Matrix{self.n}x{self.m}{self.entries!r}'

    # hooks ptina/common.py would patch on real Taichi (unused here because of _tinahacked)
    def element_wise_writeback_binary(self, *a): raise NotImplementedError
    def is_global(self): return False


def _as_scalar(a):
    """Python numbers inside vectors take Taichi's default types when they meet a binary op"""
    if isinstance(a, builtins.bool):
        return i32(builtins.int(a))
    if isinstance(a, builtins.int):
        return i32(a)
    if isinstance(a, builtins.float):
        return f32(a)
    return a


def Vector(entries, dt=None):
    return Matrix(list(entries))


class _FieldMatrix(Matrix):
    """A vector/matrix field element: reads like a value, component assignment writes through to the field."""

    def __init__(self, field, idx, entries, n, m):
        Matrix.__init__(self, _entries=entries, n=n, m=m)
        self._field, self._idx = field, idx

    def _set(self, k, val):
        self.entries[k] = val
        self._field._store(self._idx, k, val)


# ------------------------------------------------------------------------------------------------------------------
# fields
# ------------------------------------------------------------------------------------------------------------------
def _is_int_dtype(dt):
    return dt in (builtins.int, i32, u32, ti_int) or dt is np.int32


def _norm_idx(idx):
    if idx is None:
        return ()
    if isinstance(idx, Matrix):
        return tuple(builtins.int(_pyval(e)) for e in idx.entries)
    if isinstance(idx, tuple):
        out = []
        for e in idx:
            if isinstance(e, Matrix):
                out.extend(builtins.int(_pyval(x)) for x in e.entries)
            elif e is not None:
                out.append(builtins.int(_pyval(e)))
        return tuple(out)
    return (builtins.int(_pyval(idx)),)


def _to_store(val, is_int):
    v = _pyval(val)
    if is_int:
        if isinstance(v, builtins.float):
            v = _trunc_i32(v)
        return _wi(builtins.int(v))
    return builtins.float(v)


class ScalarField:
    def __init__(self, dtype, shape=None):
        self.is_int = _is_int_dtype(dtype)
        self.dtype = dtype
        self.arr = None
        if shape is not None:
            self._alloc(shape)

    def _alloc(self, shape):
        if isinstance(shape, (builtins.int, np.integer)):
            shape = (builtins.int(shape),)
        self.arr = np.zeros(tuple(builtins.int(s) for s in shape), np.int32 if self.is_int else np.float32)

    @property
    def shape(self):
        return self.arr.shape

    def __getitem__(self, idx):
        v = self.arr[_norm_idx(idx)]
        return i32(builtins.int(v)) if self.is_int else _f(builtins.float(v))

    def __setitem__(self, idx, val):
        self.arr[_norm_idx(idx)] = _to_store(val, self.is_int)

    def fill(self, v):
        self.arr[...] = v

    def from_numpy(self, a):
        a = np.asarray(a)
        if self.is_int:
            self.arr[...] = (a.astype(np.int64) & 0xFFFFFFFF).astype(np.uint32).view(np.int32).reshape(self.arr.shape)
        else:
            self.arr[...] = a.astype(np.float32).reshape(self.arr.shape)

    def to_numpy(self):
        return self.arr.copy()


class MatrixField:
    def __init__(self, n, m, dtype, shape):
        self.n, self.m, self.is_int = n, m, _is_int_dtype(dtype)
        if isinstance(shape, (builtins.int, np.integer)):
            shape = (builtins.int(shape),)
        self.fshape = tuple(builtins.int(s) for s in shape)
        self.arr = np.zeros(self.fshape + (n * m,), np.int32 if self.is_int else np.float32)

    @property
    def shape(self):
        return self.fshape

    def __getitem__(self, idx):
        idx = _norm_idx(idx)
        row = self.arr[idx]
        ent = [i32(builtins.int(v)) for v in row] if self.is_int else [_f(builtins.float(v)) for v in row]
        return _FieldMatrix(self, idx, ent, self.n, self.m)

    def _store(self, idx, k, val):
        self.arr[idx + (k,)] = _to_store(val, self.is_int)

    def __setitem__(self, idx, val):
        idx = _norm_idx(idx)
        if isinstance(val, Matrix):
            vals = val.entries
        elif isinstance(val, (list, tuple, np.ndarray)):
            vals = list(np.asarray(val, dtype=object).reshape(-1)) if not isinstance(val, np.ndarray) else list(val.reshape(-1))
            if vals and isinstance(vals[0], (list, tuple)):
                vals = [x for r in val for x in r]
        else:
            vals = [val] * (self.n * self.m)
        assert len(vals) == self.n * self.m, (len(vals), self.n, self.m)
        self.arr[idx] = [_to_store(v, self.is_int) for v in vals]

    def fill(self, v):
        self.arr[...] = v

    def from_numpy(self, a):
        self.arr[...] = np.asarray(a).astype(self.arr.dtype).reshape(self.arr.shape)

    def to_numpy(self):
        return self.arr.reshape(self.fshape + ((self.n,) if self.m == 1 else (self.n, self.m))).copy()


def field(dtype, shape=None):
    return ScalarField(dtype, shape)


Vector.field = lambda n, dtype, shape: MatrixField(n, 1, dtype, shape)
Vector.unit = Matrix.unit
Vector.zero = lambda dt, n: Matrix.zero(dt, n)
Matrix.field = staticmethod(lambda n, m, dtype, shape: MatrixField(n, m, dtype, shape))


class _SNode:
    def __init__(self, shape=()):
        self.shape = tuple(shape)

    def dense(self, axes, n):
        return _SNode(self.shape + (builtins.int(_pyval(n)),))

    def place(self, *fields):
        for f in fields:
            f._alloc(self.shape)


root = _SNode()
i, j, k, l = 'i', 'j', 'k', 'l'
ij, ijk = 'ij', 'ijk'

# ------------------------------------------------------------------------------------------------------------------
# decorators / runtime
# ------------------------------------------------------------------------------------------------------------------


def _byvalue(fn):
    def wrapped(*args, **kwargs):
        args = [a.copy() if isinstance(a, Matrix) else a for a in args]
        return fn(*args, **kwargs)
    wrapped.__name__ = getattr(fn, '__name__', 'func')
    wrapped.__wrapped__ = fn
    return wrapped


def func(fn):
    return _byvalue(fn)


pyfunc = func


def kernel(fn):
    return fn


def data_oriented(cls):
    return cls


def static(x, *xs):
    return [x] + list(xs) if xs else x


def ndrange(*dims):
    rs = []
    for d in dims:
        if isinstance(d, (tuple, list)):
            rs.append(range(builtins.int(_pyval(d[0])), builtins.int(_pyval(d[1]))))
        else:
            rs.append(range(builtins.int(_pyval(d))))
    return itertools.product(*rs)


def template():
    return None


def ext_arr():
    return None


def materialize_callback(fn):
    fn()
    return fn


def expr_init(x):
    if isinstance(x, (dict, str)) or x is None:
        return x
    if isinstance(x, Matrix):
        return x.copy()
    return x


expr_init_func = expr_init


def assign(a, b):
    raise NotImplementedError('ti.assign is only reachable through Taichi AST transforms')


def inside_kernel():
    return True


class _Runtime:
    materialized = True
    default_ip = i32
    default_fp = f32


def get_runtime():
    return _Runtime


impl = types.SimpleNamespace(get_runtime=get_runtime)
cpu, cuda, opengl, cc, metal, gpu = 'cpu', 'cuda', 'opengl', 'cc', 'metal', 'gpu'
cfg = types.SimpleNamespace(arch=cpu, cpu_max_num_threads=8)


def init(*a, **k):
    pass


def get_os_name():
    return 'linux'


class _TaichiOperations:
    pass


lang = types.SimpleNamespace(common_ops=types.SimpleNamespace(TaichiOperations=_TaichiOperations))
TaichiOperations = _TaichiOperations

# ------------------------------------------------------------------------------------------------------------------
# math
# ------------------------------------------------------------------------------------------------------------------


def _map(fn, x, *rest):
    if isinstance(x, Matrix):
        return Matrix(_entries=[fn(e, *rest) for e in x.entries], n=x.n, m=x.m)
    return fn(x, *rest)


def _ff(x):
    return builtins.float(_pyval(x)) if not isinstance(x, f32) else x.v


def sqrt(x): return _map(lambda e: _f(math.sqrt(_rf(_ff(e))) if _ff(e) >= 0 else math.nan), x)
def floor(x): return _map(lambda e: _f(math.floor(_ff(e))) if math.isfinite(_ff(e)) else _f(_ff(e)), x)
def ceil(x): return _map(lambda e: _f(math.ceil(_ff(e))) if math.isfinite(_ff(e)) else _f(_ff(e)), x)
def sin(x): return _map(lambda e: _f(_libm.sinf(_rf(_ff(e)))), x)
def cos(x): return _map(lambda e: _f(_libm.cosf(_rf(_ff(e)))), x)
def tan(x): return _map(lambda e: _f(_libm.tanf(_rf(_ff(e)))), x)
def log(x): return _map(lambda e: _f(_libm.logf(_rf(_ff(e)))), x)
def exp(x): return _map(lambda e: _f(_libm.expf(_rf(_ff(e)))), x)
def atan2(y, x): return _f(_libm.atan2f(_rf(_ff(y)), _rf(_ff(x))))
def pow(a, b): return a ** b


def cast(x, dt):
    if isinstance(x, Matrix):
        return x.cast(dt)
    if dt in (u32,):
        v = _pyval(x)
        return u32(_trunc_i32(v) if isinstance(v, builtins.float) else v)
    if _is_int_dtype(dt):
        return ti_int(x)
    return ti_float(x)


def bit_cast(x, dt):
    raise NotImplementedError


def _fminmax(a, b, want_max):
    a, b = _as_scalar(_pyval_keep(a)), _as_scalar(_pyval_keep(b))
    if isinstance(a, f32) or isinstance(b, f32):
        x, y = _ff(a), _ff(b)
        x, y = _rf(x), _rf(y)
        if x != x:
            return _f(y)
        if y != y:
            return _f(x)
        return _f(builtins.max(x, y) if want_max else builtins.min(x, y))
    r = builtins.max(a.v, b.v) if want_max else builtins.min(a.v, b.v)
    return (u32 if isinstance(a, u32) or isinstance(b, u32) else i32)._mk(r)


def _pyval_keep(x):
    if isinstance(x, np.integer):
        return builtins.int(x)
    if isinstance(x, np.floating):
        return builtins.float(x)
    return x


def _minmax(args, want_max):
    if len(args) == 1 and isinstance(args[0], (list, tuple)):
        args = tuple(args[0])
    r = args[0]
    for o in args[1:]:
        if isinstance(r, Matrix) or isinstance(o, Matrix):
            if not isinstance(r, Matrix):
                r = Matrix(_entries=[r] * len(o.entries), n=o.n, m=o.m)
            r = r._ew(o, lambda a, b: _fminmax(a, b, want_max))
        else:
            r = _fminmax(r, o, want_max)
    return r


def ti_max(*args): return _minmax(args, True)
def ti_min(*args): return _minmax(args, False)


max = ti_max
min = ti_min


def atomic_max(a, b):
    assert isinstance(a, Matrix)
    old = a.copy()
    for q in range(len(a.entries)):
        a._set(q, _fminmax(a.entries[q], b.entries[q] if isinstance(b, Matrix) else b, True))
    return old


def atomic_min(a, b):
    assert isinstance(a, Matrix)
    old = a.copy()
    for q in range(len(a.entries)):
        a._set(q, _fminmax(a.entries[q], b.entries[q] if isinstance(b, Matrix) else b, False))
    return old


def ti_abs(x):
    return abs(x) if isinstance(x, (Matrix, _Scalar)) else builtins.abs(x)


def ti_int(x=0, *a):
    """`int(...)` inside kernels: cast to i32 (truncation; u32 reinterprets).  Python numbers stay Python ints."""
    if a:
        return builtins.int(x, *a)
    if isinstance(x, Matrix):
        return Matrix(_entries=[ti_int(e) for e in x.entries], n=x.n, m=x.m)
    if isinstance(x, f32):
        return i32._mk(_trunc_i32(x.v))
    if isinstance(x, _Int):
        return i32._mk(x.v)
    return builtins.int(x)


def ti_float(x=0.0):
    if isinstance(x, Matrix):
        return Matrix(_entries=[ti_float(e) for e in x.entries], n=x.n, m=x.m)
    if isinstance(x, _Scalar):
        return _f(builtins.float(x.v))
    if isinstance(x, builtins.int) and not isinstance(x, builtins.bool):
        return _f(builtins.float(x))
    return builtins.float(x)


_rng_state = [0x2545F491]


def random(dt=None):
    # xorshift32 -- only the MLT engine calls ti.random(); Taichi's own stream cannot be reproduced (parity unpinned there)
    s = _rng_state[0]
    s ^= (s << 13) & 0xFFFFFFFF
    s ^= s >> 17
    s ^= (s << 5) & 0xFFFFFFFF
    _rng_state[0] = s
    return _f((s >> 8) / 16777216.0)


def imread(path):
    from PIL import Image as _I
    return np.asarray(_I.open(path)).swapaxes(0, 1)[:, ::-1]


KERNEL_BUILTINS = {'int': ti_int, 'float': ti_float, 'min': ti_min, 'max': ti_max, 'abs': ti_abs}
