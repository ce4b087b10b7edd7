"""TEST INFRASTRUCTURE (oracle).  Restatement of ptina/sampling/sobol.py:32-70 `calc_sobol_vgrid`.

The reference reads the Joe-Kuo `new-joe-kuo-6.21201` direction numbers as a flat integer stream
`s, a, m_1..m_s` per dimension (d >= 2) from the third-party package `pysobol`
(requirements.txt:6, unpinned, module `pysobol.data._sobol_data`), which is not installed here.
The same table ships with scipy (`scipy/stats/_sobol_direction_numbers.npz`: `poly`, `vinit`);
`joe_kuo_stream()` re-serialises it into pysobol's stream format and `calc_sobol_vgrid` below then
follows the reference line by line.  Cross-check (tests/test_sobol.py): scipy.stats.qmc.Sobol and
torch.quasirandom.SobolEngine, unscrambled.
"""
import os
import numpy as np


def joe_kuo_stream(ndims):
    """Flat [s, a, m_1..m_s] stream for dimensions 2..ndims (what `iter(pysobol.data._sobol_data)` yields)."""
    import scipy
    path = os.path.join(os.path.dirname(scipy.__file__), 'stats', '_sobol_direction_numbers.npz')
    data = np.load(path)
    poly, vinit = data['poly'], data['vinit']
    out = []
    for j in range(1, ndims):
        p = int(poly[j])
        s = p.bit_length() - 1
        a = (p >> 1) & ((1 << (s - 1)) - 1) if s > 1 else 0
        out.append(s)
        out.append(a)
        out.extend(int(x) for x in vinit[j, :s])
    return out


def calc_sobol_vgrid(N=2**20, D=21201, stream=None):
    """sobol.py:32-70, literal (numpy int64 arithmetic as in the reference)."""
    file = iter(stream if stream is not None else joe_kuo_stream(D))
    L = int(np.ceil(np.log2(N)))
    V = np.full((L + 1, D), 0)
    for j in range(D):
        if j != 0:
            s = next(file)
            a = next(file)
            m = np.full(s + 1, 0)
            for i in range(s):
                m[i + 1] = next(file)
        else:
            m = np.full(L + 1, 1)
            s = L
        if L <= s:
            for i in range(L + 1):
                V[i, j] = m[i] << (32 - i)
        else:
            for i in range(s + 1):
                V[i, j] = m[i] << (32 - i)
            for i in range(s + 1, L + 1):
                V[i, j] = V[i - s, j] ^ (V[i - s, j] >> s)
                for k in range(1, s):
                    V[i, j] ^= ((a >> (s - 1 - k)) & 1) * V[i - k, j]
    return V


_cache = {}


def vgrid_i32(D=21201):
    """The table as it lands in the i32 Taichi field (sobol.py:81-88: from_numpy wraps int64 -> i32)."""
    if D not in _cache:
        V = calc_sobol_vgrid(2**20, D)
        _cache[D] = (V & 0xFFFFFFFF).astype(np.uint32).view(np.int32).copy()
    return _cache[D]
