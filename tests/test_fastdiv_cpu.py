"""The exact integer division / modulo shortcuts of ptb_math.cuh (FastDiv, FastMod), restated with Python integers: every formula the
kernels evaluate with __umulhi / __umul64hi is checked against // and % on boundary and random inputs for the divisors that occur
(path slots per sample and tile rows of real film sizes, the Sobol dimension count 21201) and for adversarial ones."""
import random

M32, M64 = (1 << 32) - 1, (1 << 64) - 1


def make_fastdiv(d):
    L = 0
    while (1 << L) < d:
        L += 1
    if d & (d - 1) == 0:
        return dict(M=0, sh=L)
    M = ((1 << (31 + L)) // d) + 1
    assert M <= M32, (d, M)
    return dict(M=M, sh=L - 1)


def fastdiv(n, f):
    return (((n * f['M']) >> 32) >> f['sh']) if f['M'] else (n >> f['sh'])      # __umulhi(n, M) >> sh


def make_fastmod(m):
    if m <= 1:
        return dict(m=m, M=0, K=0)
    return dict(m=m, M=(M64 // m + 1) & M64, K=(1 << 32) % m)


def fastpymod(i, f):
    if f['M'] == 0:
        return 0
    n = i & M32                                   # (unsigned)i
    q = ((n * f['M']) >> 64) & M32                # (unsigned)__umul64hi(n, M)
    r = (n - q * f['m']) & M32
    r = r - (1 << 32) if r >= (1 << 31) else r    # (int)
    if i < 0:
        r -= f['K']
        if r < 0:
            r += f['m']
    return r


def _probe_values(limit, d, rng):
    vals = {0, 1, 2, limit - 1, limit - 2, d - 1, d, d + 1, 2 * d - 1, 2 * d, 3 * d + 1}
    for k in (1, 2, 3, 7, 1000, limit // d - 1, limit // d):
        for e in (-1, 0, 1):
            vals.add(k * d + e)
    vals |= {rng.randrange(limit) for _ in range(4000)}
    return [v for v in vals if 0 <= v < limit]


def test_fastdiv_exact_below_2_31():
    rng = random.Random(5)
    films = [(512, 512), (1920, 1080), (1024, 1024), (96, 96), (128, 72), (130, 70), (12, 10), (1, 1), (2048, 1024), (333, 777)]
    divisors = {3, 5, 6, 7, 9, 10, 11, 12, 13, 24, 25, 48, 127, 129, 255, 257, 1023, 1025, 65535, 65537, (1 << 20) + 1, (1 << 30) + 1, (1 << 31) - 1, (1 << 30) - 1}
    for nx, ny in films:
        ty = (ny + 3) // 4
        divisors |= {ty, ((nx + 7) // 8) * ty * 32}
    divisors |= {rng.randrange(1, 1 << 31) for _ in range(300)} | {1 << k for k in range(0, 31)}
    for d in sorted(divisors):
        f = make_fastdiv(d)
        for n in _probe_values(1 << 31, d, rng):
            assert fastdiv(n, f) == n // d, (n, d)


def test_fastmod_python_semantics_over_int32():
    rng = random.Random(6)
    for m in [1, 2, 3, 7, 21201, 21200, 21202, 65536, 65537, 1000003, (1 << 31) - 1, 1 << 20] + [rng.randrange(1, 1 << 31) for _ in range(200)]:
        f = make_fastmod(m)
        vals = {0, 1, -1, m, -m, m - 1, -(m - 1), m + 1, -(m + 1), (1 << 31) - 1, -(1 << 31), -(1 << 31) + 1, (1 << 31) - 2}
        vals |= {rng.randrange(-(1 << 31), 1 << 31) for _ in range(3000)}
        vals |= {k * m + e for k in (-3, -2, 2, 3, 101201, -101201) for e in (-1, 0, 1) if -(1 << 31) <= k * m + e < (1 << 31)}
        for i in vals:
            assert fastpymod(i, f) == i % m, (i, m)
