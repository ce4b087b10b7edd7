"""Sobol sampler state (reference: ptina/sampling/sobol.py:32-125).

The reference builds the direction-number grid V[21][21201] on the host from the Joe-Kuo `new-joe-kuo-6.21201`
table (via the `pysobol` package) and advances one Gray-code step per `render()`.  Here the same grid is built with
vectorised NumPy from the copy of that table that ships inside SciPy, uploaded once, and the device evaluates the
Gray-code point of any index k in closed form (X_k = XOR of V[b+1] over the set bits b of k ^ (k >> 1)), which is
what lets a sample range be sharded across GPUs.  `time` counts update() calls exactly like the reference.
"""
import os
import numpy as np

from ..common import Singleton
from .. import _native

DIM = 21201          # sobol.py:75
NSAMPLES = 2**20
SKIP = 64


def _joe_kuo(ndims):
    """(s, a, m[ndims, 18]) per dimension from SciPy's bundled new-joe-kuo-6.21201 table."""
    import scipy
    path = os.path.join(os.path.dirname(scipy.__file__), 'stats', '_sobol_direction_numbers.npz')
    if not os.path.exists(path):
        raise RuntimeError('Joe-Kuo direction numbers not found (scipy/stats/_sobol_direction_numbers.npz)')
    table = np.load(path)
    poly = table['poly'][:ndims].astype(np.int64)
    vinit = table['vinit'][:ndims].astype(np.int64)
    s = np.floor(np.log2(poly)).astype(np.int64)          # degree of the primitive polynomial
    a = (poly >> 1) & ((1 << np.maximum(s - 1, 0)) - 1)   # inner coefficients a_1..a_{s-1}
    return s, a, vinit


def calc_sobol_vgrid(N=NSAMPLES, D=DIM):
    """Direction numbers V[L+1][D] scaled to 32 bits, wrapped to int32 (what lands in the reference's i32 field)."""
    L = int(np.ceil(np.log2(N)))
    s, a, m = _joe_kuo(D)
    V = np.zeros((L + 1, D), dtype=np.int64)
    # dimension 0: van der Corput (all m_i = 1)
    for i in range(1, L + 1):
        V[i, 0] = 1 << (32 - i)
    cols = np.arange(1, D)
    sj, aj = s[1:], a[1:]
    for i in range(1, L + 1):
        init = i <= sj                                      # V_i = m_i * 2^(32-i) for i <= s
        mi = m[1:, min(i, m.shape[1]) - 1] if i <= m.shape[1] else np.zeros(D - 1, np.int64)
        direct = np.where(init, mi << (32 - i), 0)
        prev = V[np.maximum(i - sj, 0), cols]               # V_{i-s}
        rec = prev ^ (prev >> sj)
        for k in range(1, int(sj.max())):                   # XOR of a_k * V_{i-k}, k = 1..s-1
            use = (~init) & (k < sj) & (((aj >> np.maximum(sj - 1 - k, 0)) & 1) == 1)
            rec = np.where(use, rec ^ V[max(i - k, 0), cols], rec)
        V[i, 1:] = np.where(init, direct, rec)
    V[0, 0] = 1 << 32
    return (V & 0xFFFFFFFF).astype(np.uint32).view(np.int32)


class SobolSampler(metaclass=Singleton):
    def __init__(self, dim=DIM, nsamples=NSAMPLES, skip=SKIP):
        self.dim, self.nsamples, self.skip = dim, nsamples, skip
        self.V = calc_sobol_vgrid(nsamples, dim)
        _native.context().set_sobol_table(self.V)    # also performs reset(): time = skip

    def reset(self):
        _native.context().sobol_reset()

    def update(self):
        ctx = _native.context()
        ctx.sobol_time = ctx.sobol_time + 1

    @property
    def time(self):
        return _native.context().sobol_time

    def points(self, k=None):
        """P[dim] of Sobol point index k (default: the current one) -- the table `calc(i)` indexes (sobol.py:107-109)."""
        ctx = _native.context()
        return ctx.sobol_point(ctx.sobol_time if k is None else k)
