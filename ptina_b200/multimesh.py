"""Compose several meshes (each with its own world transform and material id) into the flat triangle
soup `ModelPool().load(vertices, mtlids)` takes.

Drop-in for the reference's host helper ptina/multimesh.py:9-87 (same name, same input/output
contract); host-side NumPy in float64, exactly as there -- this is the step before the hot path.

    primitives: iterable of (p, n, t, w, m)
        p  [num, 3, 3] vertex positions          n  [num, 3, 3] vertex normals
        t  [num, 3, 2] texcoords or None         w  [4, 4] world matrix          m  material id or None
    returns (vertices [num_total*3, 8] = px py pz nx ny nz u v,  mtlids [num_total])
"""
import numpy as np


def _transform(p, n, t, w, m):
    if w is None or p is None or n is None:
        raise AssertionError('primitive needs positions, normals and a world matrix')
    nfaces = p.shape[0]
    pos = np.asarray(p, dtype=np.float64).reshape(nfaces * 3, 3)
    nrm = np.asarray(n, dtype=np.float64).reshape(nfaces * 3, 3)
    uv = (np.zeros((nfaces * 3, 2)) if t is None else np.asarray(t, dtype=np.float64).reshape(nfaces * 3, 2))
    assert pos.shape[0] == nrm.shape[0] == uv.shape[0]
    wt = np.asarray(w, dtype=np.float64).T
    hp = np.concatenate([pos, np.ones((pos.shape[0], 1))], axis=1) @ wt       # points: w = 1
    hn = np.concatenate([nrm, np.zeros((nrm.shape[0], 1))], axis=1) @ wt      # directions: w = 0
    pos = hp[:, :3] / hp[:, 3:4]
    nrm = hn[:, :3] / np.linalg.norm(hn[:, :3], axis=1, keepdims=True)
    return np.concatenate([pos, nrm, uv], axis=1), np.full(nfaces, -1 if m is None else m)


def compose_multiple_meshes(primitives):
    parts = [_transform(*prim) for prim in primitives]
    assert len(parts), 'no primitives'
    vertices = np.concatenate([a for a, _ in parts], axis=0)
    mtlids = np.concatenate([b for _, b in parts], axis=0)
    assert len(vertices) == 3 * len(mtlids)
    return vertices, mtlids
