"""GPU parity tests: the CUDA path (through the C-ABI, via the PTina-style facade) against the CPU oracle on the same
seeded inputs.  Bit-exact: Morton codes, sorted (mc, id), child/leaf topology, node boxes, primary rays, primary hit ids /
depth / uv.  Float: BSDF eval / pdf <= 1e-5 relative, per-sample radiance, accumulated film.  Plus size-independent
properties at BASELINE.json's full sizes (ordered traversal == the reference's own traversal order, on the GPU)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle  # noqa: E402
from ptina_b200 import scenes, worker, _native  # noqa: E402

SMALL = {'cornell_boxes': (96, 96), 'cornell_monkey': (96, 96), 'matball': (96, 96), 'mega_small': (128, 72)}


def load(gpu, name, size=None, ref=True):
    sc = scenes.CONFIGS[name]()
    if size:
        sc['size'] = size
    gpu.sobol_reset()
    scenes.apply(worker, sc)
    worker.clear()
    o = None
    if ref:
        o = oracle.Oracle()
        scenes.apply(o, sc)
    return sc, o


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


@pytest.mark.parametrize('name', list(SMALL))
def test_lbvh_bit_exact(gpu, name):
    sc, o = load(gpu, name, SMALL[name])
    info = gpu.tree
    a, b = gpu.export_tree(), o.export_tree()
    for key in ('mc', 'id', 'leaf', 'child'):
        assert np.array_equal(a[key], b[key]), f'{name}: {key} differs'
    for key in ('bmin', 'bmax'):
        assert np.array_equal(bits(a[key]), bits(b[key])), f'{name}: {key} differs'
    assert info.valid == 1 and info.depth == o.validate_tree() and info.policy == _native.TRAVERSE_ORDERED
    assert info.list_n == {'cornell_boxes': 10, 'cornell_monkey': 10}.get(name, info.list_n)


@pytest.mark.parametrize('name', list(SMALL))
def test_primary_rays_and_hits_bit_exact(gpu, name):
    sc, o = load(gpu, name, SMALL[name])
    for k in (65, 66, 97, 1088):
        a, b = gpu.trace_primary(k), o.primary(k)
        assert np.array_equal(bits(a['rays']), bits(b['rays'])), f'{name} k={k}: primary rays differ'
        assert np.array_equal(a['index'], b['index']), f'{name} k={k}: {(a["index"] != b["index"]).sum()} hit ids differ'
        assert np.array_equal(a['hit'], b['hit'])
        h = b['hit'] == 1
        assert np.array_equal(bits(a['depth'])[h], bits(b['depth'])[h]) and np.array_equal(bits(a['uv'])[h], bits(b['uv'])[h])


def test_sobol_points_bit_exact(gpu):
    o = oracle.Oracle()
    for k in (1, 64, 65, 97, 4095, 1 << 19):
        assert np.array_equal(bits(gpu.sobol_point(k)), bits(o.sobol_point(k)))


@pytest.mark.parametrize('name', ['cornell_monkey', 'mega_small'])
def test_secondary_rays_match_reference_order(gpu, name):
    """Incoherent rays from random surface points: ordered traversal, the literal traversal on the GPU and the oracle agree
    on hit id, depth and uv bit for bit; shadow queries agree with `closest depth <= dis`."""
    sc, o = load(gpu, name, SMALL[name])
    rng = np.random.default_rng(3)
    m = 20000
    verts = np.asarray(sc['vertices'], np.float32)[:, :3].reshape(-1, 3, 3)
    f = rng.integers(0, verts.shape[0], m)
    w = rng.dirichlet([1, 1, 1], m).astype(np.float32)
    org = (verts[f] * w[:, :, None]).sum(1)
    d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([org, d], 1).astype(np.float32)
    avoid = f.astype(np.int32)
    ref = o.intersect(rays, avoid)
    for policy in (_native.TRAVERSE_ORDERED, _native.TRAVERSE_ORDERED_EXACT, _native.TRAVERSE_REFERENCE):
        got = gpu.intersect(rays, avoid, policy)
        assert np.array_equal(got['index'], ref['index']), f'policy {policy}: {(got["index"] != ref["index"]).sum()} ids differ'
        h = ref['hit'] == 1
        assert np.array_equal(bits(got['depth'])[h], bits(ref['depth'])[h]) and np.array_equal(bits(got['uv'])[h], bits(ref['uv'])[h])
    dis = np.where(ref['hit'] == 1, ref['depth'] * rng.choice([0.5, 1.0, 1.5], m), 5.0).astype(np.float32)
    want = ((ref['hit'] == 1) & (ref['depth'] <= dis)).astype(np.int32)
    for policy in (_native.TRAVERSE_ORDERED, _native.TRAVERSE_ORDERED_EXACT, _native.TRAVERSE_REFERENCE):
        assert np.array_equal(gpu.occluded(rays, dis, avoid, policy), want)


@pytest.mark.parametrize('name', ['cornell_boxes', 'cornell_monkey', 'matball', 'mega_small'])
def test_traversal_structure_is_a_conservative_tree(gpu, name):
    """The structure the production kernel walks (PLOC topology for small scenes, the LBVH's for big ones): a proper binary tree
    from node 0 in which every triangle that is neither listed nor untestable is a leaf exactly once, every stored child box is
    the exact union of what is below it, leaf boxes are the inflated bounds (the gate box for ill-conditioned triangles, which
    carry the no-distance-cull flag up to the root), and the height fits the traversal stack."""
    sc, o = load(gpu, name, SMALL[name], ref=False)
    t = gpu.export_traversal()
    n = t['leaf_lo'].shape[0]
    assert t['ploc']        # PLOC topology at every size (single-block kernel up to 8193 triangles, the multi-block one beyond)
    nodes = t['nodes']
    ids = np.stack([nodes[:, 3].view(np.int32), nodes[:, 7].view(np.int32)], 1)
    lo = np.stack([nodes[:, 0:3], nodes[:, 8:11]], 1); hi = np.stack([nodes[:, 4:7], nodes[:, 12:15]], 1)
    flags = t['leaf_lo'][:, 3].view(np.int32)
    seen = np.zeros(n, np.int32)
    visited = np.zeros(n - 1, bool)
    # iterative post-order: box and must-flag of every internal node from its stored child entries
    box_lo, box_hi, must_of, height = {}, {}, {}, {}
    stack = [(0, False)]
    while stack:
        j, done = stack.pop()
        if not done:
            assert not visited[j], 'node reached twice'
            visited[j] = True
            stack.append((j, True))
            for k in range(2):
                c = ids[j, k]
                if c >= 0 and (c & 0x3FFFFFFF) >= n:
                    stack.append(((c & 0x3FFFFFFF) - n, False))
            continue
        l, h, m, ht = [], [], False, 1
        for k in range(2):
            c = ids[j, k]
            if c < 0:
                continue
            cm, c = bool(c & 0x40000000), c & 0x3FFFFFFF
            if c < n:
                seen[c] += 1
                assert (flags[c] & (1 | 8)) == 0, 'a listed / untestable triangle is in the tree'
                is_must = bool(flags[c] & 2)
                want_lo, want_hi = (t['gbox'][2 * c, :3], t['gbox'][2 * c + 1, :3]) if is_must else (t['leaf_lo'][c, :3], t['leaf_hi'][c, :3])
                assert cm == is_must
            else:
                want_lo, want_hi, is_must = box_lo[c - n], box_hi[c - n], must_of[c - n]
                assert cm == is_must
                ht = max(ht, 1 + height[c - n])
            assert np.array_equal(lo[j, k], want_lo) and np.array_equal(hi[j, k], want_hi)
            l.append(lo[j, k]); h.append(hi[j, k]); m = m or cm
        assert l, 'internal node with nothing below it is referenced'
        box_lo[j], box_hi[j], must_of[j], height[j] = np.min(l, 0), np.max(h, 0), m, ht
    active = (flags & (1 | 8)) == 0
    assert np.array_equal(seen, active.astype(np.int32))
    assert height[0] <= 64
    assert np.all(ids[~visited] == -1) or not t['ploc']          # PLOC: entries it did not use are marked empty


@pytest.mark.parametrize('name', ['cornell_boxes', 'cornell_monkey', 'matball', 'mega_small'])
def test_traversal_stack_stays_within_the_tree_height(gpu, name):
    """The resident tree kernels keep the whole stack in shared memory as 16 32-bit entries when the build says the tree is shallow
    enough (wavefront.cu ptb_s16_ok: PLOC height - 1 <= 16): the deepest stack the production traversal reaches on incoherent rays
    must stay below the height the build reports (one entry per level that has two internal children)."""
    sc, o = load(gpu, name, SMALL[name])
    t = gpu.tree
    rng = np.random.default_rng(9)
    m = 40000
    verts = np.asarray(sc['vertices'], np.float32)[:, :3].reshape(-1, 3, 3)
    lo, hi = verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)
    org = rng.uniform(lo - 0.5, hi + 0.5, (m, 3))
    d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([org, d], 1).astype(np.float32)
    gpu.set_counting(True)
    try:
        gpu.reset_counters()
        gpu.intersect(rays, None, _native.TRAVERSE_ORDERED)
        got = gpu.counters()
    finally:
        gpu.set_counting(False)
    assert got['rays'] == m and got['node_visits'] > 0
    assert t.trav_depth >= 1 and got['max_stack'] <= (t.trav_depth - 1 if t.trav_ploc else t.trav_depth), (got['max_stack'], t.trav_depth)


@pytest.mark.parametrize('name', ['cornell_monkey', 'mega_small'])
def test_literal_traversal_work_counts_equal_the_oracle(gpu, name):
    """Integer parity of the work itself: under the literal policy (lbvh.py:313-347) the GPU pops, box-tests and triangle-tests
    exactly as often as the oracle on the same rays, and needs the same stack depth (incoherent secondary rays with an avoid id,
    and the primary rays of one sample)."""
    sc, o = load(gpu, name, SMALL[name])
    rng = np.random.default_rng(5)
    m = 5000
    verts = np.asarray(sc['vertices'], np.float32)[:, :3].reshape(-1, 3, 3)
    f = rng.integers(0, verts.shape[0], m)
    w = rng.dirichlet([1, 1, 1], m).astype(np.float32)
    org = (verts[f] * w[:, :, None]).sum(1)
    d = rng.normal(size=(m, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([org, d], 1).astype(np.float32)
    want = o.intersect(rays, f.astype(np.int32), counters=True)['counters']
    keys = ('node_visits', 'box_tests', 'tri_tests', 'max_stack')
    gpu.set_counting(True)
    try:
        gpu.reset_counters()
        gpu.intersect(rays, f.astype(np.int32), _native.TRAVERSE_REFERENCE)
        got = gpu.counters()
        assert got['rays'] == m and {k: got[k] for k in keys} == {k: want[k] for k in keys}
        # the primary rays of one sample (coherent, no avoid id)
        prim = o.primary(66, trace=False)['rays']
        want = o.intersect(prim, counters=True)['counters']
        gpu.reset_counters()
        gpu.intersect(prim, None, _native.TRAVERSE_REFERENCE)
        got = gpu.counters()
        assert got['rays'] == prim.shape[0] and {k: got[k] for k in keys} == {k: want[k] for k in keys}
    finally:
        gpu.set_counting(False)


@pytest.mark.parametrize('name', ['cornell_boxes', 'cornell_monkey', 'mega_small'])
def test_adversarial_rays_match_reference_order(gpu, name):
    """Rays built to stress the conservative box test + exact gate test of the production kernel: axis-parallel and nearly
    axis-parallel directions (|d| around the 1e-6 threshold of Box.intersect), rays aimed exactly at triangle vertices / edge
    midpoints / box corners (grazing the gates), origins on box planes, far-away origins, zero and non-finite directions.  Every
    policy must reproduce the oracle's literal traversal bit for bit."""
    sc, o = load(gpu, name, SMALL[name])
    rng = np.random.default_rng(17)
    tree = o.export_tree()
    verts = np.asarray(sc['vertices'], np.float32)[:, :3].reshape(-1, 3, 3)
    nf = verts.shape[0]
    rays, avoid = [], []

    def add(org, d, av=None):
        org = np.asarray(org, np.float32).reshape(-1, 3); d = np.asarray(d, np.float32).reshape(-1, 3)
        rays.append(np.concatenate([org, np.broadcast_to(d, org.shape) if d.shape[0] == 1 else d], 1))
        avoid.append(np.full(org.shape[0], -1, np.int32) if av is None else np.asarray(av, np.int32))

    m = 4000
    lo, hi = verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)
    ctr, ext = (lo + hi) / 2, (hi - lo)
    # 1. axis-parallel and almost axis-parallel directions
    for axis in range(3):
        for tiny in (0.0, 5e-7, 9.99e-7, 1e-6, 1.01e-6, 1e-5, -5e-7, -1.5e-6):
            d = rng.normal(size=(m // 8, 3)).astype(np.float32)
            d[:, axis] = tiny
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            d[:, axis] = tiny
            org = (ctr + (rng.uniform(-0.7, 0.7, (m // 8, 3)) * ext)).astype(np.float32)
            add(org, d)
    for axis in range(3):          # exactly along an axis, both signs, from inside and outside
        for sgn in (1.0, -1.0):
            d = np.zeros((m // 4, 3), np.float32); d[:, axis] = sgn
            add((ctr + rng.uniform(-0.8, 0.8, (m // 4, 3)) * ext).astype(np.float32), d)
    # 2. aimed exactly at vertices, edge midpoints and node-box corners
    f = rng.integers(0, nf, m)
    org = (ctr + rng.uniform(-0.45, 0.45, (m, 3)) * ext).astype(np.float32)
    tgt = verts[f, rng.integers(0, 3, m)]
    d = tgt - org; d /= np.linalg.norm(d, axis=1, keepdims=True)
    add(org, d)
    k = rng.integers(0, 3, m)
    tgt = (verts[f, k] + verts[f, (k + 1) % 3]) * np.float32(0.5)
    d = tgt - org; d /= np.linalg.norm(d, axis=1, keepdims=True)
    add(org, d)
    nb = tree['bmin'].shape[0]
    j = rng.integers(0, nb, m)
    pick = rng.integers(0, 2, (m, 3)).astype(bool)
    tgt = np.where(pick, tree['bmin'][j], tree['bmax'][j]).astype(np.float32)
    d = tgt - org; d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    add(org, d)
    # 3. origins exactly on box planes / triangle vertices (with and without avoid), far-away origins
    org = np.where(pick, tree['bmin'][j], tree['bmax'][j]).astype(np.float32)
    add(org, _unit(rng, m))
    add(verts[f, 0], _unit(rng, m), f)
    add(verts[f, 1], _unit(rng, m))
    far = (ctr + _unit(rng, m) * np.float32(1e4) * np.linalg.norm(ext)).astype(np.float32)
    d = (ctr + rng.uniform(-0.5, 0.5, (m, 3)) * ext - far); d /= np.linalg.norm(d, axis=1, keepdims=True)
    add(far, d)
    # 4. degenerate directions
    bad = np.array([[0, 0, 0], [np.nan, 0.3, 0.4], [np.inf, 0, 0], [1e-20, 1e-20, 1e-20], [0, 1, 0], [1e30, 1e30, 1e30], [np.nan] * 3], np.float32)
    for b in bad:
        add((ctr + rng.uniform(-0.4, 0.4, (16, 3)) * ext).astype(np.float32), b[None])
    add(np.full((4, 3), np.nan, np.float32), _unit(rng, 4))
    rays = np.ascontiguousarray(np.concatenate(rays, 0), np.float32); avoid = np.ascontiguousarray(np.concatenate(avoid, 0))
    ref = o.intersect(rays, avoid)
    h = ref['hit'] == 1
    assert h.sum() > len(h) // 10
    for policy in (_native.TRAVERSE_ORDERED, _native.TRAVERSE_ORDERED_EXACT, _native.TRAVERSE_REFERENCE):
        got = gpu.intersect(rays, avoid, policy)
        bad_i = np.nonzero(got['index'] != ref['index'])[0]
        assert bad_i.size == 0, f'policy {policy}: {bad_i.size} ids differ, first ray {rays[bad_i[0]]} got {got["index"][bad_i[0]]} want {ref["index"][bad_i[0]]}'
        assert np.array_equal(got['hit'], ref['hit'])
        assert np.array_equal(bits(got['depth'])[h], bits(ref['depth'])[h]) and np.array_equal(bits(got['uv'])[h], bits(ref['uv'])[h])
    with np.errstate(invalid='ignore'):
        dis = np.where(h, ref['depth'] * rng.choice([0.5, 1.0, 1.0, 1.5], len(h)), 5.0).astype(np.float32)
        want = (h & (ref['depth'] <= dis)).astype(np.int32)
    for policy in (_native.TRAVERSE_ORDERED, _native.TRAVERSE_ORDERED_EXACT, _native.TRAVERSE_REFERENCE):
        assert np.array_equal(gpu.occluded(rays, dis, avoid, policy), want), f'policy {policy}: shadow queries differ'


def _soup(rng, kind):
    """Triangle soups that stress the production traversal's conservative machinery: needle / sliver triangles (ill-conditioned
    Face.intersect arithmetic -> bounded by their gate box), exactly degenerate ones (never hit), triangles spanning the whole scene
    (always-test list, and its overflow), tiny ones far from the origin (large absolute rounding), duplicates (depth ties)."""
    tris = []
    def add(v0, e1, e2):
        tris.append(np.stack([v0, v0 + e1, v0 + e2], 1))
    m = 700
    c = rng.uniform(-1, 1, (m, 3)); add(c, rng.normal(size=(m, 3)) * 0.08, rng.normal(size=(m, 3)) * 0.08)            # ordinary
    c = rng.uniform(-1, 1, (150, 3)); e = rng.normal(size=(150, 3)) * 0.3
    add(c, e, e * rng.uniform(0.2, 1.5, (150, 1)) + rng.normal(size=(150, 3)) * 10.0 ** rng.uniform(-7, -2, (150, 1)))        # slivers
    c = rng.uniform(-1, 1, (40, 3)); e = rng.normal(size=(40, 3)) * 0.2
    add(c, e, e * 2.0); add(c, e, e * 0.0)                                                                        # degenerate
    c = rng.uniform(-1, 1, (60, 3)) * 0.7 + 6.0; add(c, rng.normal(size=(60, 3)) * 1e-3, rng.normal(size=(60, 3)) * 1e-3)   # tiny, off-centre
    nbig = {'few_big': 6, 'many_big': 50, 'no_big': 0}[kind]
    if nbig:
        c = rng.uniform(-1.2, -0.8, (nbig, 3)); add(c, rng.uniform(4.5, 6.5, (nbig, 3)), rng.uniform(4.5, 6.5, (nbig, 3)) * [1, -1, 1])
    t = np.concatenate(tris, 0).astype(np.float32)
    t = np.concatenate([t, t[:12]], 0)                                                                            # exact duplicates
    n = t.shape[0]
    nrm = np.cross(t[:, 1] - t[:, 0], t[:, 2] - t[:, 0]); nrm /= np.maximum(np.linalg.norm(nrm, axis=1, keepdims=True), 1e-20)
    v = np.zeros((n, 3, 8), np.float32); v[:, :, :3] = t; v[:, :, 3:6] = nrm[:, None, :]
    return v.reshape(n * 3, 8)


@pytest.mark.parametrize('kind', ['few_big', 'many_big', 'no_big'])
def test_triangle_soup_matches_reference_order(gpu, kind):
    from ptina_b200.model import ModelPool
    from ptina_b200.tree import BVHTree
    for seed in range(40):       # Morton-code runs of >= 3 make the reference's hierarchy a non-tree (it raises): take the first proper one
        rng = np.random.default_rng(1000 * seed + {'few_big': 5, 'many_big': 6, 'no_big': 7}[kind])
        verts = _soup(rng, kind)
        nf = verts.shape[0] // 3
        ModelPool().load(verts, np.zeros(nf, np.int32))
        try:
            BVHTree().build()
        except RuntimeError:
            continue
        if gpu.tree.valid:
            break
    o = oracle.Oracle(); o.load_model(verts, np.zeros(nf, np.int32)); o.build_tree()
    info = gpu.tree
    assert info.valid == 1 and info.policy == _native.TRAVERSE_ORDERED, 'no proper tree found: the production traversal was not exercised'
    a, b = gpu.export_tree(), o.export_tree()
    for key in ('mc', 'id', 'leaf', 'child'):
        assert np.array_equal(a[key], b[key])
    assert info.list_n == {'few_big': 6, 'many_big': 32, 'no_big': 0}[kind]
    tri = verts[:, :3].reshape(nf, 3, 3)
    m = 30000
    f = rng.integers(0, nf, m)
    w = rng.dirichlet([1, 1, 1], m).astype(np.float32)
    # targets: interior points, exact vertices, exact edge midpoints, points just outside an edge
    tgt = (tri[f] * w[:, :, None]).sum(1)
    sel = rng.integers(0, 4, m)
    k = rng.integers(0, 3, m)
    tgt = np.where((sel == 1)[:, None], tri[f, k], tgt)
    mid = (tri[f, k] + tri[f, (k + 1) % 3]) * np.float32(0.5)
    tgt = np.where((sel == 2)[:, None], mid, tgt)
    out = mid + (mid - tri[f, (k + 2) % 3]) * (10.0 ** rng.uniform(-7, -3, (m, 1))).astype(np.float32)
    tgt = np.where((sel == 3)[:, None], out, tgt).astype(np.float32)
    org = rng.uniform(-3, 3, (m, 3)).astype(np.float32)
    org[: m // 4] = (tri[rng.integers(0, nf, m // 4)] * rng.dirichlet([1, 1, 1], m // 4).astype(np.float32)[:, :, None]).sum(1)    # origins on surfaces
    d = tgt - org; d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    d[m // 2: m // 2 + 2000, rng.integers(0, 3)] *= np.float32(1e-4)                                                 # nearly axis-parallel
    rays = np.ascontiguousarray(np.concatenate([org, d], 1), np.float32)
    avoid = np.where(rng.random(m) < 0.3, f, -1).astype(np.int32)
    ref = o.intersect(rays, avoid)
    h = ref['hit'] == 1
    assert h.mean() > 0.3
    policies = (_native.TRAVERSE_AUTO, _native.TRAVERSE_ORDERED_EXACT, _native.TRAVERSE_REFERENCE)
    for policy in policies:
        got = gpu.intersect(rays, avoid, policy)
        bad = np.nonzero(got['index'] != ref['index'])[0]
        assert bad.size == 0, f'{kind} policy {policy}: {bad.size} ids differ, first ray {rays[bad[0]]} got {got["index"][bad[0]]} want {ref["index"][bad[0]]}'
        assert np.array_equal(bits(got['depth'])[h], bits(ref['depth'])[h]) and np.array_equal(bits(got['uv'])[h], bits(ref['uv'])[h])
    dis = np.where(h, ref['depth'] * rng.choice([0.5, 1.0, 1.0, 1.5], m), 5.0).astype(np.float32)
    want = (h & (ref['depth'] <= dis)).astype(np.int32)
    for policy in policies:
        assert np.array_equal(gpu.occluded(rays, dis, avoid, policy), want), f'{kind} policy {policy}: shadow queries differ'


def test_triangle_soup_films_identical_between_policies(gpu):
    """Whole path-traced frames over a soup with slivers, degenerate and scene-spanning triangles: the production traversal, the
    exact-ordered one and the literal one give bit-identical films (extend and shadow rays, five bounces)."""
    from ptina_b200.model import ModelPool
    from ptina_b200.tree import BVHTree
    from ptina_b200.tools import matrix as mx
    sc, _ = load(gpu, 'cornell_boxes', (96, 96), ref=False)          # camera, materials, light of the Cornell scene ...
    for seed in range(40):
        rng = np.random.default_rng(7000 + seed)
        verts = _soup(rng, 'many_big')
        verts[:, :3] = verts[:, :3] * np.float32(0.9) + np.float32([0.0, 2.0, 0.0])       # ... around a soup in front of it
        nf = verts.shape[0] // 3
        ModelPool().load(verts, rng.integers(-1, 3, nf).astype(np.int32))
        try:
            BVHTree().build()
        except RuntimeError:
            continue
        if gpu.tree.valid:
            break
    assert gpu.tree.valid == 1 and gpu.tree.list_n == 32
    films = {}
    for policy in (_native.TRAVERSE_AUTO, _native.TRAVERSE_ORDERED_EXACT, _native.TRAVERSE_REFERENCE):
        gpu.set_traversal(policy); gpu.sobol_reset(); worker.clear()
        gpu.render(_native.ENGINE_PATH, 3)
        films[policy] = gpu.get_film().copy()
    gpu.set_traversal(_native.TRAVERSE_AUTO)
    ref = films[_native.TRAVERSE_REFERENCE]
    assert np.isfinite(ref).all() and (ref[..., :3] > 0).mean() > 0.5
    for policy in (_native.TRAVERSE_AUTO, _native.TRAVERSE_ORDERED_EXACT):
        assert np.array_equal(bits(films[policy]), bits(ref)), f'policy {policy}: {(bits(films[policy]) != bits(ref)).sum()} film words differ'


def _random_materials(rng, m):
    p = np.zeros((m, 14), np.float32)
    p[:, 0:3] = rng.uniform(0.05, 1.0, (m, 3))
    p[:, 3] = rng.choice([0.0, 0.3, 1.0], m)                  # metallic
    p[:, 4] = rng.uniform(0.05, 1.0, m)                       # roughness
    p[:, 5] = rng.uniform(0.0, 1.0, m)                        # specular
    p[:, 6] = rng.uniform(0.0, 1.0, m)
    p[:, 7] = rng.choice([0.0, 0.5], m)                       # subsurface
    p[:, 8] = rng.choice([0.0, 0.7], m)                       # sheen
    p[:, 9] = rng.uniform(0.0, 1.0, m)
    p[:, 10] = rng.choice([0.0, 0.0, 1.0], m)                 # clearcoat
    p[:, 11] = rng.uniform(0.0, 1.0, m)
    p[:, 12] = rng.choice([0.0, 0.0, 0.5, 1.0], m)            # transmission
    p[:, 13] = rng.uniform(1.1, 2.0, m)                       # ior
    return p


def _unit(rng, m, upper=None):
    v = rng.normal(size=(m, 3))
    if upper is not None:
        v[:, 2] = np.abs(v[:, 2]) * upper
    return (v / np.linalg.norm(v, axis=1, keepdims=True)).astype(np.float32)


def relerr(a, b, floor=1e-6):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)


def test_bsdf_eval_within_1e5(gpu):
    rng = np.random.default_rng(11)
    m = 50000
    params = _random_materials(rng, m)
    nrm = _unit(rng, m)
    wi, wo = _unit(rng, m), _unit(rng, m)
    flip = (wi * nrm).sum(1) < 0
    wi[flip] = -wi[flip]                                         # the integrator always passes the facing normal, sign >= 0
    geom = np.concatenate([nrm, np.ones((m, 1), np.float32), wi, wo], 1)
    a, b = gpu.eval_bsdf(params, geom), oracle.eval_bsdf(params, geom)
    assert np.isfinite(b).all()
    # north_star: per-path BSDF eval within 1e-5 relative (relative to the value's own magnitude; floor for ~0 results)
    scale = np.maximum(np.abs(b).max(1, keepdims=True), 1e-4)
    assert (np.abs(a - b) / scale).max() <= 1e-5, (np.abs(a - b) / scale).max()


def test_bsdf_sample_within_1e5(gpu):
    rng = np.random.default_rng(12)
    m = 50000
    params = _random_materials(rng, m)
    params[:, 10] = 0.0                                           # clearcoat sampling is NaN-driven in the reference (sample_GTR1); below
    nrm = _unit(rng, m)
    wi = _unit(rng, m)
    flip = (wi * nrm).sum(1) < 0
    wi[flip] = -wi[flip]
    samp = rng.random((m, 3)).astype(np.float32)
    geom = np.concatenate([nrm, np.ones((m, 1), np.float32), wi, samp], 1)
    a, b = gpu.sample_bsdf(params, geom), oracle.sample_bsdf(params, geom)
    assert np.isfinite(b).all()
    assert np.abs(a[:, :3] - b[:, :3]).max() <= 2e-5                                   # unit directions: absolute
    # The sampled half vector goes through sincosf, where CUDA's libm and glibc differ by an ulp or two.  GTR2's
    # t = 1 + (alpha^2-1) cosh^2 cancels near the peak, so that ulp is amplified by ~4/t in Ds (hence in pdf and colour):
    # the 1e-5 bound of north_star is asserted where the function is well conditioned (t >= 0.1), and 1e-5 + 1e-6/t elsewhere
    # (same INPUT directions give <= 1e-5 everywhere: test_bsdf_eval_within_1e5).
    alpha2 = np.maximum(0.001, params[:, 4] ** 2) ** 2
    h = wi + b[:, :3]
    h = h / np.maximum(np.linalg.norm(h, axis=1, keepdims=True), 1e-20)
    cosh = np.clip((h * nrm).sum(1), 0, 1)
    t = 1 + (alpha2 - 1) * cosh ** 2
    refracted = (b[:, :3] * nrm).sum(1) < 0
    t = np.where(refracted, alpha2, np.maximum(t, alpha2))
    tol = 1e-5 + 1e-6 / t
    e_pdf = relerr(a[:, 3], b[:, 3], 1e-4)
    cs = np.maximum(np.abs(b[:, 4:]).max(1), 1e-4)
    e_col = np.abs(a[:, 4:] - b[:, 4:]).max(1) / cs
    assert (e_pdf <= tol).all() and (e_col <= tol).all(), (float((e_pdf / tol).max()), float((e_col / tol).max()))
    well = t >= 0.1
    assert well.mean() > 0.5 and e_pdf[well].max() <= 1e-5 and e_col[well].max() <= 1e-5
    # every lobe exercised
    assert (b[:, 3] == 0).any() and np.isclose(b[:, 3], 1 / np.pi).any() and (params[:, 12] > 0).any()


def _adversarial_bsdf_inputs(rng, m):
    """Materials with exact zeros / odd values and direction sets built to hit the guards of the production BSDF forms."""
    params = _random_materials(rng, m)
    k = m // 8
    params[:k, 12] = 0.0; params[:k, 10] = 0.0; params[:k, 7] = 0.0                       # the config 1 / 2 / 4 kind: everything eliminable
    params[k:2 * k, 12] = -0.0; params[k:2 * k, 10] = -0.0                                   # negative zeros
    odd = slice(2 * k, 3 * k)
    params[odd, 13] = rng.choice([0.0, 1e-4, 1.0, 1e5, np.inf, np.nan], k)                  # ior outside the `sane` range
    params[odd, 4] = rng.choice([0.0, 1e-3, 1.0, 50.0, 2e4], k)                              # roughness
    params[odd, 11] = rng.choice([0.0, 1.0, 1.0101, 2.0, -5.0, 1e6], k)                      # clearcoatGloss: clearcoatAlpha 0 / negative / 1
    params[odd, 0] = rng.choice([0.0, 1.0, 1e5, np.nan], k)
    params[odd, 12] = 0.0; params[odd, 10] = 0.0; params[odd, 7] = 0.0
    nrm = _unit(rng, m)
    wi, wo = _unit(rng, m), _unit(rng, m)
    sign = np.ones((m, 1), np.float32)
    j = np.arange(m)
    wo[j % 7 == 0] = -wi[j % 7 == 0]                                                         # half vector 0 / 0
    r = j % 7 == 1                                                                           # mirror direction: half vector == normal (cosh = 1)
    wo[r] = 2 * (wi[r] * nrm[r]).sum(1, keepdims=True) * nrm[r] - wi[r]
    wo[j % 7 == 2] *= np.float32(1e3)                                                        # not unit: the magnitude guards
    wi[j % 7 == 3] *= np.float32(1e20)
    g = j % 7 == 4                                                                           # grazing: cosi + coso ~ 0
    t = np.cross(nrm[g], _unit(rng, int(g.sum())))
    t /= np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-20)
    wi[g] = t.astype(np.float32); wo[g] = -t.astype(np.float32) + np.float32(1e-7) * nrm[g]
    sign[j % 5 == 0] = -1.0                                                                  # the taps accept it; the integrator never passes it
    return params, nrm, sign, wi, wo


def test_bsdf_production_forms_equal_the_literal_ones_bitwise(gpu):
    # ptb_shade.cuh "terms that vanish exactly": the production disney_brdf / disney_bounce leave out factors an exactly-zero weight
    # annihilates (under guards that make the skipped factor finite) and share the tail of the three lobes; the literal forms evaluate
    # disney.py:52-233 term by term.  Same bits (a zero may differ in sign), NaNs in the same places.
    rng = np.random.default_rng(2026)
    m = 280000
    params, nrm, sign, wi, wo = _adversarial_bsdf_inputs(rng, m)
    geom = np.concatenate([nrm, sign, wi, wo], 1).astype(np.float32)
    with np.errstate(all='ignore'):
        a, b = gpu.eval_bsdf(params, geom), gpu.eval_bsdf_literal(params, geom)
        assert np.array_equal(a, b, equal_nan=True), int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())
        assert np.isnan(b).any() and np.isinf(b).any() and (b == 0).any()                    # the fall-backs were exercised
        samp = rng.random((m, 3)).astype(np.float32)
        samp[::11, 2] = 0.0; samp[1::11, 2] = -0.25; samp[2::11, 0] = 1.0; samp[3::11, 0] = 0.0; samp[4::11, 2] = np.nan
        geom = np.concatenate([nrm, sign, wi, samp], 1).astype(np.float32)
        a, b = gpu.sample_bsdf(params, geom), gpu.sample_bsdf_literal(params, geom)
        assert np.array_equal(a, b, equal_nan=True), int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())
        lobes = (np.isclose(b[:, 3], 1 / np.pi).any(), (b[:, 3] == 0).any(), ((b[:, :3] * nrm).sum(1) < -0.1).any())
        assert all(lobes), lobes
    # and the plain production call on the renderer's own kind of input (eliminable material, unit directions, facing normal)
    n2 = 100000
    p2 = _random_materials(rng, n2); p2[:, 12] = 0.0; p2[:, 10] = 0.0; p2[:, 7] = 0.0
    nr, w1, w2 = _unit(rng, n2), _unit(rng, n2), _unit(rng, n2)
    flip = (w1 * nr).sum(1) < 0
    w1[flip] = -w1[flip]
    g2 = np.concatenate([nr, np.ones((n2, 1), np.float32), w1, w2], 1)
    a, b = gpu.eval_bsdf(p2, g2), gpu.eval_bsdf_literal(p2, g2)
    assert np.isfinite(b).all() and np.array_equal(a, b)
    assert (np.abs(a - oracle.eval_bsdf(p2, g2)) / np.maximum(np.abs(b).max(1, keepdims=True), 1e-4)).max() <= 1e-5


def test_clearcoat_lobe_behaves_like_reference(gpu):
    # microfacet.py:68-71: sqrt(alpha**(2-2u) - 1) is NaN for alpha < 1 -> cosoh = max(0, NaN) = 0 -> invalid sample (zeros)
    rng = np.random.default_rng(13)
    m = 4096
    params = _random_materials(rng, m)
    params[:, 10] = 1.0
    nrm = np.tile(np.array([0, 0, 1], np.float32), (m, 1))
    wi = _unit(rng, m, upper=1.0)
    samp = rng.random((m, 3)).astype(np.float32)
    samp[:, 2] *= 0.13                                             # coatrate = lerp(0.04, 0.1, 1) = 0.136 -> coat lobe
    geom = np.concatenate([nrm, np.ones((m, 1), np.float32), wi, samp], 1)
    a, b = gpu.sample_bsdf(params, geom), oracle.sample_bsdf(params, geom)
    assert np.array_equal(np.isnan(a), np.isnan(b))
    ok = ~np.isnan(b).any(1)
    assert np.allclose(a[ok], b[ok], rtol=1e-5, atol=1e-6)
    assert (b[ok][:, 3] == 0).all()


def test_material_fetch_with_textures(gpu):
    sc, o = load(gpu, 'matball', SMALL['matball'])
    rng = np.random.default_rng(5)
    m = 20000
    mtlid = rng.integers(-1, 4, m).astype(np.int32)
    uv = rng.uniform(-1.5, 2.5, (m, 2)).astype(np.float32)          # exercises the wrap mode (Python-style %)
    a, b = gpu.material_get(mtlid, uv), o.material_get(mtlid, uv)
    assert relerr(a, b, 1e-3).max() <= 1e-5


def test_lights_and_world(gpu):
    sc, o = load(gpu, 'cornell_boxes', SMALL['cornell_boxes'])
    # add a POINT light as well so both kinds are exercised; order matters for LightPool.hit (first hit wins)
    from ptina_b200.tools import matrix as mx
    for api in (worker, o):
        api.add_light(mx.translate((0.5, 2.0, 0.5)), np.array([3.0, 2.0, 1.0]), 0.3, 'POINT')
    rng = np.random.default_rng(6)
    m = 20000
    org = rng.uniform([-1.9, 0.1, -1.9], [1.9, 3.9, 1.9], (m, 3))
    rays = np.concatenate([org, _unit(rng, m)], 1).astype(np.float32)
    a, b = gpu.light_hit(rays), o.light_hit(rays)
    assert np.array_equal(a[:, 0], b[:, 0]) and (b[:, 0] == 1).sum() > 100
    assert relerr(a[:, 1:], b[:, 1:], 1e-4).max() <= 1e-5
    hs = np.concatenate([org, rng.random((m, 3))], 1).astype(np.float32)
    a, b = gpu.light_sample(hs), o.light_sample(hs)
    assert relerr(a[:, [0, 4]], b[:, [0, 4]], 1e-4).max() <= 1e-5            # distance, pdf
    assert np.abs(a[:, 1:4] - b[:, 1:4]).max() <= 2e-6                         # unit direction (POINT lights go through sincosf)
    assert (np.abs(a[:, 5:] - b[:, 5:]) / np.maximum(np.abs(b[:, 5:]).max(1, keepdims=True), 1e-4)).max() <= 2e-5
    sc, o = load(gpu, 'matball', SMALL['matball'])
    d = _unit(rng, m)
    a, b = gpu.world_at(d), o.world_at(d)
    # texel lookups near a bilinear cell edge can land on the neighbouring cell (atan2 differs by an ulp): compare robustly
    bad = relerr(a, b, 1e-3).max(1) > 1e-4
    assert bad.mean() < 2e-3, bad.mean()


@pytest.mark.parametrize('name,engine', [('cornell_boxes', 'path'), ('cornell_monkey', 'path'), ('matball', 'brute'), ('mega_small', 'path')])
def test_single_sample_radiance(gpu, name, engine):
    sc, o = load(gpu, name, SMALL[name])
    ge = {'path': _native.ENGINE_PATH, 'brute': _native.ENGINE_BRUTE}[engine]
    oe = {'path': oracle.ENGINE_PATH, 'brute': oracle.ENGINE_BRUTE}[engine]
    for k in (65, 80):
        a, b = gpu.render_sample(ge, k), o.render_sample(oe, k)
        assert not np.isnan(a).any()
        rel = (np.abs(a - b) / np.maximum(np.abs(b).max(2, keepdims=True), 1e-2)).max(2)
        # libm differences (sin/cos/atan2/pow, <= 2 ulp) occasionally flip a discrete decision deep in a path; those pixels
        # are isolated.  Everything else agrees to float rounding.
        assert (rel > 1e-4).mean() < 2e-3, f'{name} k={k}: {(rel > 1e-4).mean():.2e} of pixels differ'
        assert np.median(rel) < 1e-6


@pytest.mark.parametrize('name,engine,spp', [('cornell_boxes', 'path', 32), ('cornell_monkey', 'path', 32), ('matball', 'brute', 32), ('mega_small', 'path', 16)])
def test_accumulated_film_rel_rmse(gpu, name, engine, spp):
    """Same Sobol indices on both sides: per-pixel relative RMSE of the accumulated image <= 1e-3 (north_star's gate is stated
    for converged 1024-spp images; with identical sample sets it already holds at a few spp)."""
    sc, o = load(gpu, name, SMALL[name])
    ge = {'path': _native.ENGINE_PATH, 'brute': _native.ENGINE_BRUTE}[engine]
    oe = {'path': oracle.ENGINE_PATH, 'brute': oracle.ENGINE_BRUTE}[engine]
    gpu.render(ge, spp)
    o.render(oe, spp)
    a, b = worker.get_image(), o.get_image()
    assert a.shape == b.shape == (*SMALL[name], 4) and np.all(a[..., 3] == 1)
    fa, fb = gpu.get_film(), o.get_film()
    assert np.all(fa[..., 3] == spp) and np.all(fb[..., 3] == spp)
    rmse = np.sqrt((((a[..., :3] - b[..., :3]) / np.maximum(b[..., :3], 1e-2)) ** 2).mean())
    assert rmse <= 1e-3, rmse
    out_a, out_b = np.zeros(a.shape[0] * a.shape[1] * 3, np.float32), np.zeros(a.shape[0] * a.shape[1] * 3, np.float32)
    worker.fast_export_image(out_a); o.fast_export_image(out_b)
    assert np.array_equal(out_a.reshape(a.shape[1], a.shape[0], 3), a[..., :3].transpose(1, 0, 2))
    assert gpu.sobol_time == 64 + spp


def test_batching_and_ranges_are_equivalent(gpu):
    """render(n) == n x render(1) == rank-interleaved render_range shards summed (what the multi-GPU path does)."""
    sc, o = load(gpu, 'cornell_monkey', (64, 64), ref=False)
    gpu.render(_native.ENGINE_PATH, 6)
    a = gpu.get_film().copy()
    gpu.sobol_reset(); worker.clear()
    for _ in range(6):
        gpu.render(_native.ENGINE_PATH, 1)
    b = gpu.get_film().copy()
    assert np.array_equal(bits(a), bits(b))
    worker.clear()
    from ptina_b200.dist import shard_range
    for rank in range(4):
        first, count, stride = shard_range(65, 6, rank, 4)
        gpu.render_range(_native.ENGINE_PATH, first, count, stride)
    c = gpu.get_film()
    assert np.allclose(a, c, rtol=1e-5, atol=1e-6) and np.array_equal(a[..., 3], c[..., 3])


@pytest.mark.parametrize('name', ['cornell_monkey', 'mega_small'])
def test_overlapped_schedule_is_bit_identical_to_serial(gpu, name):
    """The production schedule runs the shadow stage of a bounce beside the extend stage of the next one (two streams); with
    per-stage timing on, the stages run one after the other.  Same additions in the same order per path: identical films."""
    sc, o = load(gpu, name, SMALL[name], ref=False)
    gpu.render(_native.ENGINE_PATH, 8)
    a = gpu.get_film().copy()
    gpu.sobol_reset(); worker.clear()
    gpu.set_counting(False, True)            # profiling: serial schedule
    try:
        gpu.render(_native.ENGINE_PATH, 8)
        b = gpu.get_film().copy()
    finally:
        gpu.set_counting(False, False)
    assert np.array_equal(bits(a), bits(b)) and a[..., 3].max() == 8


def test_full_size_ordered_equals_reference_order(gpu):
    """BASELINE configs 1/2 at full 512x512: hit ids / depth / uv of the optimised traversal equal the literal reference
    traversal run on the GPU, for primary rays of several Sobol points (size-independent property)."""
    for name in ('cornell_boxes', 'cornell_monkey'):
        sc, _ = load(gpu, name, ref=False)
        for k in (66, 97):
            gpu.set_traversal(_native.TRAVERSE_ORDERED)
            a = gpu.trace_primary(k)
            gpu.set_traversal(_native.TRAVERSE_REFERENCE)
            b = gpu.trace_primary(k)
            gpu.set_traversal(_native.TRAVERSE_AUTO)
            assert np.array_equal(a['index'], b['index']) and np.array_equal(bits(a['depth']), bits(b['depth'])) and np.array_equal(bits(a['uv']), bits(b['uv']))
        # and the whole path tracer: film from the ordered policy == film from the literal policy
        gpu.set_traversal(_native.TRAVERSE_ORDERED); gpu.sobol_reset(); worker.clear(); gpu.render(_native.ENGINE_PATH, 2); fa = gpu.get_film().copy()
        gpu.set_traversal(_native.TRAVERSE_REFERENCE); gpu.sobol_reset(); worker.clear(); gpu.render(_native.ENGINE_PATH, 2); fb = gpu.get_film().copy()
        gpu.set_traversal(_native.TRAVERSE_AUTO)
        assert np.array_equal(bits(fa), bits(fb)), f'{name}: {(bits(fa) != bits(fb)).sum()} film words differ between traversal policies'


def test_mega_scene_full_size(gpu):
    """Config 4 (~1M triangles, 1920x1080): tree is a valid Karras tree within the 32-entry stack, Morton codes sorted, and the
    optimised traversal equals the literal one on the GPU for every primary ray and for a full path-traced sample."""
    sc, _ = load(gpu, 'mega', ref=False)
    info = gpu.tree
    assert info.n == len(sc['mtlids']) > 1_000_000 and info.valid == 1 and info.depth <= 31
    t = gpu.export_tree()
    assert (np.diff(t['mc']) >= 0).all() and np.array_equal(np.sort(t['id']), np.arange(info.n))
    assert np.array_equal(t['leaf'], t['id'])
    pos = np.asarray(sc['vertices'], np.float32)[:, :3]
    assert np.array_equal(t['bmin'][0], pos.min(0)) and np.array_equal(t['bmax'][0], pos.max(0))
    gpu.set_traversal(_native.TRAVERSE_ORDERED); a = gpu.trace_primary(65)
    gpu.set_traversal(_native.TRAVERSE_REFERENCE); b = gpu.trace_primary(65)
    assert np.array_equal(a['index'], b['index']) and np.array_equal(bits(a['depth']), bits(b['depth'])) and np.array_equal(bits(a['uv']), bits(b['uv']))
    assert (a['hit'] == 1).mean() > 0.5
    gpu.set_traversal(_native.TRAVERSE_ORDERED); sa = gpu.render_sample(_native.ENGINE_PATH, 65)
    gpu.set_traversal(_native.TRAVERSE_REFERENCE); sb = gpu.render_sample(_native.ENGINE_PATH, 65)
    gpu.set_traversal(_native.TRAVERSE_AUTO)
    assert np.array_equal(bits(sa), bits(sb))


def test_edge_cases_and_errors(gpu):
    from ptina_b200.model import ModelPool
    from ptina_b200.tree import BVHTree
    sc, _ = load(gpu, 'cornell_boxes', (32, 32), ref=False)
    # stale tree is refused, not silently used
    ModelPool().load(sc['vertices'][:30], sc['mtlids'][:10])
    with pytest.raises(_native.NativeError, match='stale'):
        gpu.render(_native.ENGINE_PATH, 1)
    BVHTree().build()
    gpu.render(_native.ENGINE_PATH, 1)
    # default material (mtlid None -> -1) and float64 input
    ModelPool().load(np.asarray(sc['vertices'], np.float64))
    BVHTree().build()
    o = oracle.Oracle(); scenes.apply(o, dict(sc, mtlids=-np.ones(len(sc['mtlids']), np.int32)))
    a, b = gpu.render_sample(_native.ENGINE_PATH, 70), o.render_sample(oracle.ENGINE_PATH, 70)
    assert np.median(np.abs(a - b)) < 1e-6
    # two triangles: smallest real tree
    ModelPool().load(sc['vertices'][:6], sc['mtlids'][:2]); BVHTree().build()
    assert gpu.tree.valid == 1 and gpu.tree.depth == 2
    o = oracle.Oracle(); scenes.apply(o, dict(sc, vertices=sc['vertices'][:6], mtlids=sc['mtlids'][:2]))
    assert np.array_equal(gpu.trace_primary(65)['index'], o.primary(65)['index'])
    # capacity errors keep the reference's messages
    with pytest.raises(AssertionError, match='too many faces'):
        ModelPool().load(np.zeros((3 * 2**21, 8), np.float32))
    with pytest.raises(_native.NativeError, match='exceeds max_filmsize'):
        gpu.set_size(2048, 2048)
    from ptina_b200.image import ImagePool
    with pytest.raises(RuntimeError, match='Out of memory!'):
        ImagePool().load([np.zeros((2048, 2049), np.float32)])
    # duplicate-heavy geometry (coincident triangles -> runs of equal Morton codes): the reference's tree is not a proper tree;
    # the library must notice, fall back to the reference's traversal order, and still agree with the oracle
    v = np.tile(np.asarray(sc['vertices'][:3], np.float32), (9, 1)); v[3 * 4:3 * 5, :3] += 0.5
    for api in (worker, ):
        ModelPool().load(v, np.zeros(9, np.int32)); BVHTree().build()
    o = oracle.Oracle(); scenes.apply(o, dict(sc, vertices=v, mtlids=np.zeros(9, np.int32)))
    ta, tb = gpu.export_tree(), o.export_tree()
    assert np.array_equal(ta['child'], tb['child']) and np.array_equal(ta['mc'], tb['mc'])
    if o.validate_tree() < 0:
        assert gpu.tree.valid == 0 and gpu.tree.policy == _native.TRAVERSE_REFERENCE
    assert np.array_equal(gpu.trace_primary(65)['index'], o.primary(65)['index'])
    # leaf slots travel in 24-bit fields of the ray queues (PTB_SLOT_MASK): a context sized for more faces is refused when it is created
    with pytest.raises(_native.NativeError, match='2\\^24'):
        _native.Context(max_faces=(1 << 24) + 1)
    load(gpu, 'cornell_boxes', (32, 32), ref=False)


def test_preview_engine_passes(gpu):
    sc, o = load(gpu, 'cornell_monkey', (64, 64))
    from ptina_b200.engine import PreviewEngine
    PreviewEngine().render(2)
    alb, nrm = worker.get_image(1), worker.get_image(2)
    prim = o.primary(65)
    hit = prim['hit'].reshape(64, 64) == 1
    assert np.all(alb[..., 3] == 1) and np.all(nrm[..., 3] == 1)
    ln = np.linalg.norm(nrm[hit][:, :3], axis=1)                                   # mean of two unit normals (silhouette pixels may mix)
    assert (np.abs(ln - 1) < 0.05).mean() > 0.97
    assert (alb[hit][:, :3].max(1) > 0).mean() > 0.97
    assert np.allclose(worker.get_image(0), [0.9, 0.4, 0.9, 0.0])                  # pass 0 untouched -> magenta


def test_native_library_is_what_ran(gpu, native_so):
    import os
    maps = open('/proc/self/maps').read()
    assert os.path.basename(os.environ.get('PTINA_B200_LIB', 'libptina_b200.so')) in maps
    assert gpu.launches() > 0


def test_exact_division_shortcut(gpu):
    """div_exact(a, d, RN(1/d)) -- the 5-FMA quotient the slab and barycentric tests use -- equals the IEEE quotient a / d bit for
    bit on 2^30 operand pairs in the slab test's ranges and 2^28 over [1e-18, 1e18] (checked on the device)."""
    assert gpu.selftest(0, 1 << 30, seed=1) == 0
    assert gpu.selftest(1, 1 << 28, seed=2) == 0


def test_mlt_engine(gpu):
    """Config 5 (engine/mltpath.py): per-chain parity of the proposal's radiance against the oracle's path_trace driven by the same
    32-D sample vector (RNGProxy), the accept/reject bookkeeping, and statistical agreement of the MLT image with the path tracer.
    ti.random() cannot be reproduced (Philox here), so the chain trajectories themselves are 'parity unpinned'."""
    from ptina_b200.engine import MLTPathEngine
    sc, o = load(gpu, 'cornell_monkey', (64, 64))
    nch = 1 << 14
    eng = MLTPathEngine(nchains=nch, seed=7)
    eng.LSP[None] = 0.25; eng.Sigma[None] = 0.01
    eng.reset()
    worker.clear()
    eng.render(1)
    st = gpu.mlt_state(nch)
    X = st['X_new']
    assert X.min() >= 0.0 and X.max() < 1.0
    ref = o.trace_from_samples(X)
    rel = np.abs(st['L_new'] - ref).max(1) / np.maximum(np.abs(ref).max(1), 1e-2)
    assert np.median(rel) < 1e-6 and (rel > 1e-4).mean() < 5e-3, (float(np.median(rel)), float((rel > 1e-4).mean()))
    # first iteration: L_old = 0 -> accept = min(1, (avg(L_new)+1e-10)/1e-10) = 1 for every chain (mltpath.py:76-81)
    assert np.array_equal(st['X_old'], X) and np.array_equal(st['L_old'], st['L_new'])
    film = gpu.get_film()
    assert film[..., 3].sum() == nch                       # one splat of weight 1 per chain (mltpath.py:47-52)
    # large steps ~ LSP of the chains on the second iteration; small steps move by ~Sigma
    eng.render(1)
    st2 = gpu.mlt_state(nch)
    moved = np.abs(st2['X_new'] - st['X_old'])
    moved = np.minimum(moved, 1 - moved)                    # distance on the unit torus
    small = moved.max(1) < 0.1
    assert 0.70 < small.mean() < 0.80, small.mean()
    assert 0.006 < moved[small].std() < 0.014
    # accept/reject (mltpath.py:76-81): accept = min(1, (avg L_new + 1e-10) / (avg L_old + 1e-10)); accepted <=> X_old := X_new
    a_new, a_old = st2['L_new'].mean(1) + 1e-10, st['L_old'].mean(1) + 1e-10
    accept = np.minimum(1.0, a_new / a_old)
    accepted = (st2['X_old'] == st2['X_new']).all(1)
    assert accepted[accept >= 1.0].all()                                   # never rejects an uphill move
    assert not accepted[a_new <= 1e-9].any() or (a_old <= 1e-9).any()       # a black proposal is (almost) never accepted
    mid = (accept > 0.05) & (accept < 0.95)
    assert mid.sum() > 500 and abs(accepted[mid].mean() - accept[mid].mean()) < 0.05
    assert np.array_equal(st2['L_old'][accepted], st2['L_new'][accepted]) and np.array_equal(st2['L_old'][~accepted], st['L_old'][~accepted])
    # MLT image vs path-traced image: the reference splats every PROPOSAL with weight 1 (mltpath.py:47-52, 73-74), so its image is
    # the per-pixel mean of proposed radiance -- brighter than the path tracer's mean, not an unbiased estimate; same order of magnitude
    for _ in range(40):
        eng.render(1)
    mlt = worker.get_image()[..., :3]
    worker.clear(); gpu.sobol_reset()
    gpu.render(_native.ENGINE_PATH, 64)
    pt = worker.get_image()[..., :3]
    valid = gpu.get_film()[..., 3] > 0
    ratio = mlt[valid].mean() / pt[valid].mean()
    assert 0.8 < ratio < 2.5, ratio


def test_mlt_chain_shards_reproduce_the_population(gpu):
    """chains are sharded contiguously across GPUs (dist.shard_chains) and chain c draws from the Philox stream
    keyed (seed, c, iteration) wherever it runs.  On one GPU: the two halves of a population, run as shards (first, count), leave
    exactly the per-chain state the whole population leaves, and their films add up to its film (weights exactly, colours up to the
    order of the float atomics)."""
    from ptina_b200.engine import MLTPathEngine
    load(gpu, 'cornell_monkey', (64, 64), ref=False)
    nch, half = 1 << 13, 1 << 12
    MLTPathEngine._forget()
    eng = MLTPathEngine(nchains=nch, seed=11)
    try:
        def run(first, count):
            eng.chains = (first, count)
            eng.reset()
            worker.clear()
            for _ in range(3):
                eng.render(1)
            return gpu.mlt_state(count), gpu.get_film().copy()
        full, f_full = run(0, nch)
        a, f_a = run(0, half)
        b, f_b = run(half, nch - half)
        for key in ('X_new', 'L_new', 'X_old', 'L_old'):
            assert np.array_equal(full[key][:half], a[key]) and np.array_equal(full[key][half:], b[key]), key
        assert f_full[..., 3].sum() == 3 * nch and np.array_equal(f_full[..., 3], f_a[..., 3] + f_b[..., 3])
        assert np.allclose(f_full[..., :3], f_a[..., :3] + f_b[..., :3], rtol=1e-5, atol=1e-6)
    finally:
        MLTPathEngine._forget()
