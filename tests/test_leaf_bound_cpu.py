"""Numerical check of the leaf bound the production traversal relies on (lbvh.cu k_tri_prep,
ptb_traverse.cuh trav_delta): whenever Face.intersect (geometries.py:117-148), evaluated in IEEE f32 in the reference's
operation order, ACCEPTS a ray, the real point ro + r*rd lies within eps_T + delta of the triangle's bounds.  NumPy float32
arithmetic is the same IEEE arithmetic per operation, so this is the same computation the GPU kernels and the oracle perform.
The rays are aimed at and around the edges (distances 1e-9 ... 1e-2 of the triangle's size) where acceptance is decided by the
last bits, at grazing angles down to 1e-4 rad, for triangles up to the conditioning limit."""
import numpy as np

F = np.float32
U = 2.0 ** -24


def dot32(a, b):
    return (a[:, 0] * b[:, 0] + a[:, 1] * b[:, 1]) + a[:, 2] * b[:, 2]          # left-to-right, like Taichi's .dot


def face_intersect_f32(v0, v1, v2, ro, rd):
    u, v = v1 - v0, v2 - v0
    n = np.stack([u[:, 1] * v[:, 2] - u[:, 2] * v[:, 1], u[:, 2] * v[:, 0] - u[:, 0] * v[:, 2], u[:, 0] * v[:, 1] - u[:, 1] * v[:, 0]], 1)
    b = dot32(n, rd)
    with np.errstate(all='ignore'):
        a = -dot32(n, ro - v0)
        r = a / b
        ip = ro + r[:, None] * rd
        uu, uv, vv = dot32(u, u), dot32(u, v), dot32(v, v)
        w = ip - v0
        wu, wv = dot32(w, u), dot32(w, v)
        D = uv * uv - uu * vv
        s = (uv * wv - vv * wu) / D
        t = (uv * wu - uu * wv) / D
        hit = (np.abs(b) >= F(1e-6)) & (r > 0) & (0 <= s) & (s <= 1) & (0 <= t) & (s + t <= 1)
    return hit, r


def eps_T(v0, v1, v2):
    """lbvh.cu k_tri_prep (eligible triangles)."""
    u, v = (v1 - v0).astype(np.float64), (v2 - v0).astype(np.float64)
    uu, uv, vv = (u * u).sum(1), (u * v).sum(1), (v * v).sum(1)
    D = uv * uv - uu * vv
    lu, lv = np.sqrt(uu), np.sqrt(vv)
    cond = uu * vv / np.abs(D)
    q = np.maximum(lu, lv) / np.minimum(lu, lv)
    eligible = cond * (36 + 20 * q) <= 1e6
    eps = 1.01 * (256 * U * cond * (lu + lv) + 64 * U * (np.abs(v0).max(1) + lu + lv))
    return eps, eligible


def test_accepted_points_lie_within_the_inflated_bounds():
    rng = np.random.default_rng(2026)
    worst = 0.0
    accepted = 0
    for rounds in range(6):
        m = 400000
        # triangles: scale 1e-3..1e1, offset up to 100x the size, apex angle down to the conditioning limit
        size = (10.0 ** rng.uniform(-3, 1, (m, 1))).astype(F)
        v0 = (rng.uniform(-1, 1, (m, 3)) * size * 10.0 ** rng.uniform(-1, 2, (m, 1))).astype(F)
        e1 = rng.normal(size=(m, 3)); e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
        e2 = rng.normal(size=(m, 3)); e2 -= (e2 * e1).sum(1, keepdims=True) * e1; e2 /= np.linalg.norm(e2, axis=1, keepdims=True)
        ang = 10.0 ** rng.uniform(-2.2, 0.2, (m, 1))                      # radians between the two edges
        ratio = 10.0 ** rng.uniform(-1.5, 1.5, (m, 1))
        v1 = (v0 + e1 * size).astype(F)
        v2 = (v0 + (np.cos(ang) * e1 + np.sin(ang) * e2) * size * ratio).astype(F)
        eps, eligible = eps_T(v0, v1, v2)
        # target: a point on an edge, pushed out / in by 1e-9 .. 1e-2 of the size; ray at a grazing angle down to 1e-4
        k = rng.integers(0, 3, m)
        V = np.stack([v0, v1, v2], 1).astype(np.float64)
        a, b, c = V[np.arange(m), k], V[np.arange(m), (k + 1) % 3], V[np.arange(m), (k + 2) % 3]
        lam = rng.random((m, 1))
        edge_pt = a + (b - a) * lam
        outward = edge_pt - c; outward /= np.linalg.norm(outward, axis=1, keepdims=True)
        tgt = edge_pt + outward * size * (10.0 ** rng.uniform(-9, -2, (m, 1))) * rng.choice([-1.0, 1.0], (m, 1))
        nrm = np.cross(V[:, 1] - V[:, 0], V[:, 2] - V[:, 0]); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        tang = rng.normal(size=(m, 3)); tang -= (tang * nrm).sum(1, keepdims=True) * nrm; tang /= np.linalg.norm(tang, axis=1, keepdims=True)
        graze = 10.0 ** rng.uniform(-4, 0, (m, 1))
        d = tang * np.cos(graze) + nrm * np.sin(graze) * rng.choice([-1.0, 1.0], (m, 1))
        dist = size * 10.0 ** rng.uniform(-2, 2, (m, 1))
        ro = (tgt - d * dist).astype(F)
        rd = d.astype(F); rd = (rd * (F(1) / np.sqrt(dot32(rd, rd)))[:, None]).astype(F)      # normalized() like the reference
        hit, r = face_intersect_f32(v0, v1, v2, ro, rd)
        ok = hit & eligible
        accepted += int(ok.sum())
        X = ro.astype(np.float64) + r.astype(np.float64)[:, None] * rd.astype(np.float64)       # the real point at the computed depth
        lo, hi = V.min(1), V.max(1)
        outside = np.maximum(np.maximum(lo - X, X - hi), 0).max(1)
        S = np.abs(V).max((1, 2))
        delta = 65 * U * (2 * np.abs(ro.astype(np.float64)).sum(1) + 3 * S)
        bound = eps + delta
        worst = max(worst, float((outside[ok] / bound[ok]).max()))
        assert (outside[ok] <= bound[ok]).all(), f'accepted point {outside[ok].max():.3e} outside the inflated bounds'
    assert accepted > 200000            # the sampling does straddle the acceptance boundary
    assert worst < 0.5, worst           # the bound has a safety factor of at least 2 on this sample
