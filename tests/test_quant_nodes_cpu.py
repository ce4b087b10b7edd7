"""Numerical check of the quantised traversal nodes (ptb_traverse.cuh Node32 / ray_quant / slab_quant, lbvh.cu k_quant_nodes):

1. the 16-bit field 0x8000 | q, placed by one PRMT under 0x3F000000, IS the f32 v = 1 + q / 32768;
2. the packer's planes enclose the box they quantise (lo' <= lo, hi' >= hi as real numbers) on grids built like the host does;
3. the folded slab arithmetic  x = fma(v, A, B),  A = qext * r,  B = fma(qbase, r, nc -/+ delta * r) widened by 4u|B|, brackets the
   exact plane distances of the decoded box: x_near <= (plane_near - o) / d and x_far >= (plane_far - o) / d for every axis and
   both signs of d -- the property that makes a quantised node at least as conservative as the f32 node it replaces.

NumPy float32 is the same IEEE arithmetic per operation; fma is emulated through float64 (the product of two f32 is exact there;
the extra rounding of the sum is far below the margins under test)."""
import numpy as np

F = np.float32
U = 2.0 ** -24


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def make_grid(lo, hi):
    """lbvh.cu ptb_build_tree: ext a power of two, base rounded down, widened until [lo, hi] is covered."""
    ext = 2.0 ** np.ceil(np.log2(max(float(hi) - float(lo), 1e-30) * 1.01))
    for _ in range(8):
        b = float(lo) - ext
        base = F(b)
        if float(base) > b:
            base = np.nextafter(base, F(-np.inf))
        if float(base) + ext <= float(lo) and float(base) + (2.0 - 1.0 / 32768.0) * ext >= float(hi):
            return base, F(ext)
        ext *= 2.0
    raise AssertionError('no grid')


def quant_plane(x, base, ext, up):
    """lbvh.cu quant_plane, vectorised: candidate from the estimate, corrected against the exact decoded plane."""
    b, e, xd = float(base), float(ext), x.astype(np.float64)
    t = ((xd - b) / e - 1.0) * 32768.0
    q = np.clip(np.ceil(t) if up else np.floor(t), 0, 32767).astype(np.int64)
    for _ in range(4):
        pl = b + (1.0 + q / 32768.0) * e
        bad = (pl < xd) if up else (pl > xd)
        q = np.clip(q + np.where(bad, 1 if up else -1, 0), 0, 32767)
    return q


def decode(q):
    """The kernel's PRMT: bytes (0x00, field lo, field hi, 0x3F) -> f32."""
    field = (0x8000 | q).astype(np.uint32)
    return ((np.uint32(0x3F000000) | (field << np.uint32(8))).astype(np.uint32)).view(F)


def test_prmt_field_is_the_float():
    q = np.arange(32768, dtype=np.int64)
    v = decode(q)
    assert np.array_equal(v, (1.0 + q / 32768.0).astype(F))
    assert np.all((1.0 + q / 32768.0) == v.astype(np.float64))          # exactly representable


def test_packer_encloses_boxes():
    rng = np.random.default_rng(1)
    for centre, size in ((0.0, 2.0), (5.0, 20.0), (1000.0, 3.0), (-0.3, 1e-3), (1e5, 50.0)):
        lo_s, hi_s = F(centre - size / 2), F(centre + size / 2)
        base, ext = make_grid(lo_s, hi_s)
        x = rng.uniform(float(lo_s), float(hi_s), 20000).astype(F)
        x[:4] = [lo_s, hi_s, lo_s, hi_s]
        w = (rng.uniform(0, 1, 20000) ** 4 * size).astype(F)
        lo, hi = x, np.minimum(x + w, hi_s)
        ql, qh = quant_plane(lo, base, ext, False), quant_plane(hi, base, ext, True)
        pl = float(base) + decode(ql).astype(np.float64) * float(ext)
        ph = float(base) + decode(qh).astype(np.float64) * float(ext)
        assert np.all(pl <= lo.astype(np.float64)) and np.all(ph >= hi.astype(np.float64))
        # and not looser than one grid step (plus the f32 spacing of the coordinate itself)
        step = float(ext) / 32768.0
        assert np.all(lo.astype(np.float64) - pl <= step * 1.0001) and np.all(ph - hi.astype(np.float64) <= step * 1.0001)


def test_folded_slab_brackets_the_decoded_planes():
    rng = np.random.default_rng(2)
    m = 200000
    scene_abs = F(10.0)
    base, ext = make_grid(F(-10.0), F(10.0))
    o = rng.uniform(-10, 10, m).astype(F)
    o[: m // 10] = rng.uniform(-300, 300, m // 10).astype(F)                 # cameras far outside the scene
    mag = 10.0 ** rng.uniform(-5.9, 0, m)
    d = (mag * rng.choice([-1.0, 1.0], m)).astype(F)                          # |d| from the 1e-6 threshold of Box.intersect up to 1
    ql = rng.integers(0, 32768, m)
    qh = np.minimum(ql + rng.integers(0, 2000, m), 32767)
    vl, vh = decode(ql), decode(qh)
    # per-ray constants as ray_cons / trav_delta / ray_trav / ray_quant compute them (one axis)
    r = (F(1.0) / d).astype(F)
    nc = (-(o * r)).astype(F)
    o1 = np.abs(o) * F(3.0)                                                    # |o|_1 of a 3-vector with equal components: worst case
    delta = (F(65.0) * F(U) * (F(2.0) * o1 + F(3.0) * scene_abs)).astype(F)
    nc1, nc2 = fma(-delta, r, nc), fma(delta, r, nc)
    A = (ext * r).astype(F)
    assert np.all(A.astype(np.float64) == float(ext) * r.astype(np.float64))  # exact: ext is a power of two
    k4 = np.copysign(F(4 * U), r).astype(F)
    b1 = fma(np.full(m, base, F), r, nc1); b2 = fma(np.full(m, base, F), r, nc2)
    B1, B2 = fma(-k4, np.abs(b1), b1), fma(k4, np.abs(b2), b2)
    x1, x2 = fma(vl, A, B1), fma(vh, A, B2)
    # exact distances of the decoded planes along the real ray
    od, dd = o.astype(np.float64), d.astype(np.float64)
    t1 = (float(base) + vl.astype(np.float64) * float(ext) - od) / dd
    t2 = (float(base) + vh.astype(np.float64) * float(ext) - od) / dd
    # slab_quant takes the entry / exit plane by the sign of r instead of min / max of the two distances: the same values
    assert np.array_equal(np.where(r > 0, x1, x2), np.minimum(x1, x2)) and np.array_equal(np.where(r > 0, x2, x1), np.maximum(x1, x2))
    near_c, far_c = np.minimum(x1, x2).astype(np.float64), np.maximum(x1, x2).astype(np.float64)
    near_t, far_t = np.minimum(t1, t2), np.maximum(t1, t2)
    assert np.all(near_c <= near_t) and np.all(far_c >= far_t)
    # the widening stays local: a few hundred ulps of the plane distance scale, not a scene-sized slab
    scale = (np.abs(od) + 3 * float(scene_abs) + 2 * float(ext)) * np.abs(r.astype(np.float64))
    assert np.all(near_t - near_c <= 600 * U * scale) and np.all(far_c - far_t <= 600 * U * scale)
