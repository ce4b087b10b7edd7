"""Wavefront OBJ reader for `ModelPool().load(path)` (same return contract as the reference's host loader
ptina/tools/readobj.py:21-104: dict with 'v' [nv,3], 'vt' [nt,2], 'vn' [nn,3] float32 and 'f' [nf,3,3] int32 holding
v/vt/vn indices per corner; polygons are fan-triangulated).  Host-side NumPy; this is the step before the hot path."""
import numpy as np


def _fan(corners):
    return [[corners[0], corners[k], corners[k + 1]] for k in range(1, len(corners) - 1)]


def readobj(path, scale=None):
    v, vt, vn, faces = [], [], [], []
    opener = path if hasattr(path, 'read') else open(path, 'rb')
    with opener as fh:
        for raw in fh:
            parts = raw.split()
            if not parts:
                continue
            tag = parts[0]
            if tag == b'v':
                v.append([float(x) for x in parts[1:4]])
            elif tag == b'vt':
                vt.append([float(x) for x in parts[1:3]])
            elif tag == b'vn':
                vn.append([float(x) for x in parts[1:4]])
            elif tag == b'f':
                corners = []
                for field in parts[1:]:
                    idx = [int(t) - 1 if t else 0 for t in field.split(b'/')]
                    corners.append((idx + [0, 0])[:3])
                faces.extend(_fan(corners))
    out = {
        'v': np.array(v or [[0, 0, 0]], dtype=np.float32),
        'vt': np.array(vt or [[0, 0]], dtype=np.float32),
        'vn': np.array(vn or [[0, 0, 0]], dtype=np.float32),
        'f': np.array(faces or np.zeros((1, 3, 3)), dtype=np.int32),
    }
    if scale is not None:
        out['v'] *= scale
    return out
