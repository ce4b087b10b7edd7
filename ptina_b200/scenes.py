"""Synthetic workloads for the BASELINE.json configs (the reference's `assets/` are git-ignored and do
not exist).  Pure host-side NumPy; every scene is a plain dict that `apply()` pushes
through the same loader calls a PTina driver script makes (exams/benchmark.py:7-26):

    load_materials -> load_images -> load_model -> build_tree -> clear_lights/add_light ->
    set_world_light -> set_camera -> set_size

`api` may be the `ptina_b200.worker` module or the test oracle object; both expose these names.
All generators are deterministic (fixed seeds).
"""
import numpy as np

from .multimesh import compose_multiple_meshes
from .tools import matrix as mx

# exams/benchmark.py:18-23 -- the camera of configs 1/2 (fov 60, eye ~(0, 1.945, 5.372) looking down -z)
BENCHMARK_PERS = np.array([
    [1.73205081e+00, 0.00000000e+00, 0.00000000e+00, 1.01348227e-02],
    [0.00000000e+00, 1.73205081e+00, -1.73205081e-05, -3.36860025e+00],
    [0.00000000e+00, -1.00020002e-05, -1.00020002e+00, 5.27350023e+00],
    [0.00000000e+00, -1.00000000e-05, -1.00000000e+00, 5.37243564e+00],
])

SLOTS = ('basecolor', 'metallic', 'roughness', 'specular', 'specularTint', 'subsurface', 'sheen', 'sheenTint',
         'clearcoat', 'clearcoatGloss', 'transmission', 'ior')


def material(basecolor=(0.8, 0.8, 0.8), metallic=0.0, roughness=0.4, specular=0.5, specularTint=0.4, subsurface=0.0,
             sheen=0.0, sheenTint=0.4, clearcoat=0.0, clearcoatGloss=0.5, transmission=0.0, ior=1.45, tex=None):
    """Explicit 12-slot (fac, tex) list in MaterialPool.load order (mtllib.py:58-77); tex = {slot: image id}."""
    vals = dict(basecolor=list(basecolor), metallic=metallic, roughness=roughness, specular=specular,
                specularTint=specularTint, subsurface=subsurface, sheen=sheen, sheenTint=sheenTint, clearcoat=clearcoat,
                clearcoatGloss=clearcoatGloss, transmission=transmission, ior=ior)
    tex = tex or {}
    return [(vals[s], int(tex.get(s, -1))) for s in SLOTS]


# ------------------------------------------------------------------------------------------------
# mesh primitives: each returns (p [n,3,3], nrm [n,3,3], uv [n,3,2])
# ------------------------------------------------------------------------------------------------
def quad(a, b, c, d):
    """Two triangles (a,b,c), (c,d,a) with the flat normal (b-a)x(c-a) on every corner."""
    a, b, c, d = (np.asarray(v, dtype=np.float64) for v in (a, b, c, d))
    n = np.cross(b - a, c - a)
    n = n / np.linalg.norm(n)
    p = np.array([[a, b, c], [c, d, a]])
    uv = np.array([[[0, 0], [1, 0], [1, 1]], [[1, 1], [0, 1], [0, 0]]], dtype=np.float64)
    return p, np.broadcast_to(n, p.shape).copy(), uv


def box(lo, hi):
    """Axis-aligned box, 12 triangles, outward flat normals."""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    faces = [
        ((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)),  # +z
        ((x1, y0, z0), (x0, y0, z0), (x0, y1, z0), (x1, y1, z0)),  # -z
        ((x1, y0, z1), (x1, y0, z0), (x1, y1, z0), (x1, y1, z1)),  # +x
        ((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0)),  # -x
        ((x0, y1, z1), (x1, y1, z1), (x1, y1, z0), (x0, y1, z0)),  # +y
        ((x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)),  # -y
    ]
    parts = [quad(*f) for f in faces]
    return tuple(np.concatenate([q[k] for q in parts], axis=0) for k in range(3))


def _smooth_normals(verts, faces):
    fn = np.cross(verts[faces[:, 1]] - verts[faces[:, 0]], verts[faces[:, 2]] - verts[faces[:, 0]])
    vn = np.zeros_like(verts)
    for k in range(3):
        np.add.at(vn, faces[:, k], fn)
    ln = np.linalg.norm(vn, axis=1, keepdims=True)
    return vn / np.where(ln > 0, ln, 1)


def uvsphere(segments=32, rings=16, radius=1.0, displace=None):
    """UV sphere, 2*segments*(rings-1) triangles, smooth normals; `displace(dir[n,3]) -> r[n]` scales the radius."""
    lat = np.linspace(0, np.pi, rings + 1)[1:-1]
    lon = np.linspace(0, 2 * np.pi, segments, endpoint=False)
    dirs = [np.array([[0, 1.0, 0]])]
    uvs = [np.array([[0.5, 1.0]])]
    for i, th in enumerate(lat):
        dirs.append(np.stack([np.sin(th) * np.cos(lon), np.full_like(lon, np.cos(th)), np.sin(th) * np.sin(lon)], axis=1))
        uvs.append(np.stack([lon / (2 * np.pi), np.full_like(lon, 1 - th / np.pi)], axis=1))
    dirs.append(np.array([[0, -1.0, 0]]))
    uvs.append(np.array([[0.5, 0.0]]))
    dirs, uvs = np.concatenate(dirs), np.concatenate(uvs)
    r = radius * (displace(dirs) if displace is not None else np.ones(len(dirs)))
    verts = dirs * r[:, None]
    S, R = segments, rings
    faces = []
    ring = lambda i, j: 1 + i * S + (j % S)
    for j in range(S):
        faces.append((0, ring(0, j + 1), ring(0, j)))
    for i in range(R - 2):
        for j in range(S):
            a, b, c, d = ring(i, j), ring(i, j + 1), ring(i + 1, j + 1), ring(i + 1, j)
            faces.append((a, b, c))
            faces.append((c, d, a))
    last = len(verts) - 1
    for j in range(S):
        faces.append((last, ring(R - 2, j), ring(R - 2, j + 1)))
    faces = np.array(faces)
    assert len(faces) == 2 * S * (R - 1)
    vn = _smooth_normals(verts, faces) if displace is not None else dirs
    return verts[faces], vn[faces], uvs[faces]


def icosphere(level=3, radius=1.0):
    """Subdivided icosahedron, 20*4^level near-uniform triangles, smooth (radial) normals."""
    t = (1 + 5 ** 0.5) / 2
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
         (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6),
         (7, 1, 8), (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7),
         (9, 8, 1)]
    v = [np.array(x, dtype=np.float64) / np.linalg.norm(x) for x in v]
    for _ in range(level):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = v[a] + v[b]
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    v, f = np.array(v), np.array(f)
    uv = np.stack([np.arctan2(v[:, 2], v[:, 0]) / (2 * np.pi) + 0.5, np.arccos(np.clip(v[:, 1], -1, 1)) / np.pi], axis=1)
    return (v * radius)[f], v[f], uv[f]


def room(lo=(-2.0, 0.0, -2.0), hi=(2.0, 4.0, 2.0)):
    """Cornell room, open towards +z: floor, ceiling, back, left, right -> 5 quads, inward normals.
    returns [(p, n, t)] in that order."""
    x0, y0, z0 = lo
    x1, y1, z1 = hi
    return [
        quad((x0, y0, z1), (x1, y0, z1), (x1, y0, z0), (x0, y0, z0)),   # floor  (+y)
        quad((x0, y1, z0), (x1, y1, z0), (x1, y1, z1), (x0, y1, z1)),   # ceiling(-y)
        quad((x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0)),   # back   (+z)
        quad((x0, y0, z1), (x0, y0, z0), (x0, y1, z0), (x0, y1, z1)),   # left   (+x)
        quad((x1, y0, z0), (x1, y0, z1), (x1, y1, z1), (x1, y1, z0)),   # right  (-x)
    ]


def area_light_down(center, size, color):
    """AREA light shining towards -y: LightPool uses axes@(0,0,1) as the *back* normal
    (light/__init__.py:101-113), so local +z must map to world +y."""
    world = mx.affine(np.array([[1.0, 0, 0], [0, 0, 1], [0, -1, 0]]), center)
    return (world, np.asarray(color, dtype=np.float64), float(size), 'AREA')


_WHITE = material((0.8, 0.8, 0.8), roughness=0.8)
_RED = material((0.8, 0.08, 0.06), roughness=0.8)
_GREEN = material((0.1, 0.7, 0.12), roughness=0.8)


def _cornell_base():
    prims = []
    for k, (p, n, t) in enumerate(room()):
        prims.append((p, n, t, np.identity(4), {3: 1, 4: 2}.get(k, 0)))
    lights = [area_light_down((0.0, 3.96, 0.0), 0.7, (17.0, 14.0, 10.0))]
    return prims, [_WHITE, _RED, _GREEN], lights


def cornell_boxes(nx=512, ny=512, spp=32):
    """Config 1 (exams/benchmark.py, README.md:36-44): Cornell box with two boxes, 10 + 24 = 34 triangles."""
    prims, mats, lights = _cornell_base()
    mats = mats + [material((0.75, 0.75, 0.75), roughness=0.5), material((0.7, 0.7, 0.8), roughness=0.25, metallic=0.2)]
    short = mx.translate((0.65, 0.0, 0.6)) @ mx.eularXYZ((0, -0.3, 0))
    tall = mx.translate((-0.65, 0.0, -0.55)) @ mx.eularXYZ((0, 0.35, 0))
    prims.append((*box((-0.6, 0.0, -0.6), (0.6, 1.2, 0.6)), short, 3))
    prims.append((*box((-0.6, 0.0, -0.6), (0.6, 2.4, 0.6)), tall, 4))
    v, m = compose_multiple_meshes(prims)
    assert len(m) == 34
    return dict(name='cornell_boxes', vertices=v, mtlids=m.astype(np.int32), materials=mats, images=[], lights=lights,
                world_light=([0.1] * 4, -1), pers=BENCHMARK_PERS, size=(nx, ny), engine='path', spp=spp)


def _blob_displace(dirs):
    x, y, z = dirs.T
    return (1.0 + 0.18 * np.sin(3.1 * x + 0.7) * np.cos(2.3 * y - 0.2) + 0.12 * np.sin(4.7 * z + 1.3 * x)
            + 0.25 * np.exp(-8 * ((x - 0.55) ** 2 + (y - 0.35) ** 2)) + 0.25 * np.exp(-8 * ((x + 0.55) ** 2 + (y - 0.35) ** 2)))


def cornell_monkey(nx=512, ny=512, spp=32):
    """Config 2 (README.md:46-51): the 10-triangle room + a 968-triangle closed smooth-shaded blob standing in
    for Blender's Suzanne (968 triangles) -> 978 triangles, LBVH + MIS area light."""
    prims, mats, lights = _cornell_base()
    mats = mats + [material((0.85, 0.6, 0.25), metallic=0.35, roughness=0.3, specular=0.6)]
    p, n, t = uvsphere(22, 23, 1.0, _blob_displace)
    assert p.shape[0] == 968
    world = mx.translate((0.05, 1.45, 0.1)) @ mx.eularXYZ((0.15, 0.5, 0.1))
    prims.append((p, n, t, world, 3))
    v, m = compose_multiple_meshes(prims)
    assert len(m) == 978
    return dict(name='cornell_monkey', vertices=v, mtlids=m.astype(np.int32), materials=mats, images=[], lights=lights,
                world_light=([0.1] * 4, -1), pers=BENCHMARK_PERS, size=(nx, ny), engine='path', spp=spp)


def _checker(n=512, cells=16):
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing='ij')
    c = ((i * cells // n + j * cells // n) % 2).astype(np.float32)
    img = np.stack([0.9 - 0.7 * c, 0.85 - 0.55 * c, 0.2 + 0.6 * c], axis=2).astype(np.float32)
    return img


def _envmap(nx=1024, ny=512):
    u = (np.arange(nx) + 0.5) / nx
    v = (np.arange(ny) + 0.5) / ny
    U, Vv = np.meshgrid(u, v, indexing='ij')
    sky = np.stack([0.35 + 0.4 * Vv, 0.5 + 0.35 * Vv, 0.75 + 0.25 * Vv], axis=2)
    sun = np.exp(-((U - 0.3) ** 2 * 4 + (Vv - 0.72) ** 2) * 900.0)[..., None] * np.array([60.0, 50.0, 35.0])
    ground = np.array([0.25, 0.22, 0.2]) * (Vv < 0.5)[..., None]
    img = np.where((Vv < 0.5)[..., None], ground, sky) + sun
    return img.astype(np.float32)


def matball(nx=1024, ny=1024, spp=256):
    """Config 3 (exams/matball.py with BruteEngine): glass / metal / checker-textured balls (32x16 UV spheres, 960
    triangles each) on a ground quad under an equirect environment map."""
    sp = uvsphere(32, 16, 1.0)
    mats = [
        material((0.7, 0.7, 0.7), roughness=0.6),                                     # ground
        material((0.95, 0.98, 1.0), roughness=0.05, transmission=1.0, ior=1.45),      # glass
        material((0.95, 0.75, 0.35), metallic=1.0, roughness=0.2),                     # metal
        material((1.0, 1.0, 1.0), roughness=0.5, tex={'basecolor': 0}),                # checker
    ]
    g = quad((-6, 0, 6), (6, 0, 6), (6, 0, -6), (-6, 0, -6))
    prims = [(*g, np.identity(4), 0),
             (*sp, mx.translate((-2.3, 1.0, 0.0)), 1),
             (*sp, mx.translate((0.0, 1.0, 0.0)), 2),
             (*sp, mx.translate((2.3, 1.0, 0.0)) @ mx.eularXYZ((0.3, 0.4, 0.0)), 3)]
    v, m = compose_multiple_meshes(prims)
    pers = mx.perspective(fov=45, aspect=nx / ny) @ mx.lookat(pos=(0, 0.9, 0), back=(0, 2.2, 8.0))
    return dict(name='matball', vertices=v, mtlids=m.astype(np.int32), materials=mats, images=[_checker(), _envmap()],
                lights=[], world_light=([1.0] * 4, 1), pers=pers, size=(nx, ny), engine='brute', spp=spp)


def mega(nx=1920, ny=1080, spp=1024, grid=(10, 8, 10), level=3, seed=20261018):
    """Config 4: ~1M-triangle procedural scene.  grid[0]*grid[1]*grid[2] icospheres (20*4^level triangles each,
    near-uniform so that Morton cells hold <3 centroids and the reference's LBVH stays a valid tree)
    on a jittered lattice inside an open room with one area light."""
    rng = np.random.default_rng(seed)
    gx, gy, gz = grid
    cell = 1.0
    ext = np.array([gx, gy, gz], dtype=np.float64) * cell
    prims = []
    lo, hi = -0.6 * cell * np.ones(3), ext + 0.6 * cell
    for k, (p, n, t) in enumerate(room(lo, hi)):
        prims.append((p, n, t, np.identity(4), {3: 1, 4: 2}.get(k, 0)))
    base = icosphere(level, 1.0)
    nm = 8
    mats = [_WHITE, _RED, _GREEN]
    for i in range(nm):
        hue = rng.uniform(0.15, 0.95, 3)
        mats.append(material(tuple(hue), metallic=float(rng.choice([0.0, 0.0, 0.6, 1.0])), roughness=float(rng.uniform(0.15, 0.8))))
    for ix in range(gx):
        for iy in range(gy):
            for iz in range(gz):
                r = rng.uniform(0.28, 0.42) * cell
                c = (np.array([ix, iy, iz]) + 0.5) * cell + rng.uniform(-1, 1, 3) * (0.48 * cell - r)
                w = mx.translate(c) @ mx.eularXYZ(rng.uniform(0, 6.28, 3)) @ mx.scale(r)
                prims.append((*base, w, 3 + int(rng.integers(nm))))
    v, m = compose_multiple_meshes(prims)
    center = (lo + hi) / 2
    lights = [area_light_down((center[0], hi[1] - 0.02 * cell, center[2]), 0.35 * gx * cell, (9.0, 8.5, 8.0))]
    dist = 0.5 * (hi[1] - lo[1]) / np.tan(np.radians(22.0)) * 1.05
    pers = mx.perspective(fov=44, aspect=nx / ny, near=0.05, far=500) @ mx.lookat(pos=tuple(center), back=(0, 0.0, dist + (hi[2] - center[2])))
    return dict(name='mega', vertices=v, mtlids=m.astype(np.int32), materials=mats, images=[], lights=lights,
                world_light=([0.3] * 4, -1), pers=pers, size=(nx, ny), engine='path', spp=spp)


def mega_small(nx=256, ny=144, spp=4):
    """A reduced config-4 (4x3x4 level-2 icospheres, ~15k triangles) for parity tests."""
    s = mega(nx, ny, spp, grid=(4, 3, 4), level=2)
    s['name'] = 'mega_small'
    return s


def metropolis(nx=512, ny=512):
    """Config 5 (exams/metropolis.py): the config-2 scene under MLTPathEngine."""
    s = cornell_monkey(nx, ny, spp=1)
    s.update(name='metropolis', engine='mlt')
    return s



def mini_matball():
    """Golden-fixture scene (tests/golden): a small cousin of config 3: glass / metal / textured UV spheres (8x6 -> 80 triangles each) on a ground quad, 16x16 checker,
    32x16 environment map -- sized for the Python interpreter."""
    sp = uvsphere(8, 6, 1.0)
    mats = [material((0.7, 0.7, 0.7), roughness=0.6),
            material((0.95, 0.98, 1.0), roughness=0.1, transmission=1.0, ior=1.45),
            material((0.95, 0.75, 0.35), metallic=1.0, roughness=0.25),
            material((1.0, 1.0, 1.0), roughness=0.5, tex={'basecolor': 0}),
            material((0.3, 0.5, 0.9), roughness=0.4, clearcoat=1.0, sheen=0.5, subsurface=0.3, transmission=0.5)]
    g = quad((-6, 0, 6), (6, 0, 6), (6, 0, -6), (-6, 0, -6))
    prims = [(*g, np.identity(4), 0), (*sp, mx.translate((-2.3, 1.0, 0.0)), 1), (*sp, mx.translate((0.0, 1.0, 0.0)), 2),
             (*sp, mx.translate((2.3, 1.0, 0.0)) @ mx.eularXYZ((0.3, 0.4, 0.0)), 3), (*sp, mx.translate((0.0, 1.0, 2.2)) @ mx.scale(0.6), 4)]
    v, m = compose_multiple_meshes(prims)
    pers = mx.perspective(fov=45, aspect=1) @ mx.lookat(pos=(0, 0.9, 0), back=(0, 2.2, 8.0))
    chk = _checker(16, 4)
    env = _envmap(32, 16)
    return dict(name='mini_matball', vertices=v, mtlids=m.astype(np.int32), materials=mats, images=[chk, env],
                lights=[(mx.translate((1.0, 4.0, 2.0)), np.array([30.0, 28.0, 25.0]), 0.4, 'POINT')],
                world_light=([1.0] * 4, 1), pers=pers, size=(12, 12), engine='brute', spp=2)



CONFIGS = {'mini_matball': mini_matball, 'cornell_boxes': cornell_boxes, 'cornell_monkey': cornell_monkey, 'matball': matball, 'mega': mega,
           'mega_small': mega_small, 'metropolis': metropolis}


def apply(api, scene, build=True):
    """Push a scene through the loader calls of a PTina driver (exams/benchmark.py:7-26, blender.py:555-582)."""
    api.load_materials(scene['materials'])
    api.load_images(scene['images'])
    api.load_model(scene['vertices'], scene['mtlids'])
    if build:
        api.build_tree()
    api.clear_lights()
    for world, color, size, type in scene['lights']:
        api.add_light(world, color, size, type)
    api.set_world_light(*scene['world_light'])
    api.set_camera(scene['pers'])
    api.set_size(*scene['size'])
