"""Stand-ins for the reference's git-ignored `assets/` (none of them exist), so that its driver scripts find
the files they open: exams/benchmark.py:13 and benchtiles.py:13 `assets/monkey_cornell.gltf`, exams/metropolis.py:15
`assets/cornell.gltf`, exams/matball.py:17 `assets/uvsphere.obj`.  Geometry comes from the same deterministic generators as the
BASELINE configs (ptina_b200/scenes.py); materials carry what glTF's pbrMetallicRoughness can (base colour, metallic, roughness --
readgltf hands MaterialPool three slots, the other nine stay at the table's zeros exactly as with a Blender export).

    python -m ptina_b200.tools.make_assets [directory]      (default: ./assets)
"""
import base64
import json
import os
import sys

import numpy as np

from .. import scenes


def gltf_document(vertices, mtlids, materials, name):
    """One mesh, one primitive per material id, non-indexed float32 attributes in a single embedded buffer."""
    v = np.asarray(vertices, np.float32).reshape(-1, 3, 8)
    mtlids = np.asarray(mtlids)
    blob, views, accessors, prims = b'', [], [], []

    def add(arr, kind, bounds=False):
        nonlocal blob
        arr = np.ascontiguousarray(arr, np.float32)
        views.append({'buffer': 0, 'byteOffset': len(blob), 'byteLength': arr.nbytes})
        acc = {'bufferView': len(views) - 1, 'componentType': 5126, 'count': int(arr.shape[0]), 'type': kind}
        if bounds:                                                                 # POSITION accessors must carry them
            acc['min'], acc['max'] = arr.min(0).tolist(), arr.max(0).tolist()
        accessors.append(acc)
        blob += arr.tobytes()
        return len(accessors) - 1

    for m in sorted(set(mtlids.tolist())):
        tri = v[mtlids == m].reshape(-1, 8)
        prim = {'attributes': {'POSITION': add(tri[:, 0:3], 'VEC3', True), 'NORMAL': add(tri[:, 3:6], 'VEC3'), 'TEXCOORD_0': add(tri[:, 6:8], 'VEC2')}, 'mode': 4}
        if m >= 0:
            prim['material'] = int(m)
        prims.append(prim)
    mats = []
    for k, slots in enumerate(materials):
        base, metallic, roughness = slots[0][0], slots[1][0], slots[2][0]
        mats.append({'name': f'{name}_material_{k}', 'pbrMetallicRoughness': {'baseColorFactor': [float(x) for x in (list(base) + [1.0])[:4]],
                                                                          'metallicFactor': float(metallic), 'roughnessFactor': float(roughness)}})
    return {'asset': {'version': '2.0', 'generator': 'ptina_b200.tools.make_assets'}, 'scene': 0, 'scenes': [{'name': name, 'nodes': [0]}],
            'nodes': [{'name': name, 'mesh': 0}], 'meshes': [{'name': name, 'primitives': prims}], 'materials': mats,
            'buffers': [{'byteLength': len(blob), 'uri': 'data:application/octet-stream;base64,' + base64.b64encode(blob).decode('ascii')}],
            'bufferViews': views, 'accessors': accessors}


def obj_text(p, n, t):
    """Non-indexed OBJ with v/vt/vn per corner."""
    p, n, t = (np.asarray(a, np.float32).reshape(-1, a.shape[-1]) for a in (p, n, t))
    lines = ['# ptina_b200.tools.make_assets: 32x16 UV sphere (960 triangles)']
    lines += ['v %.7g %.7g %.7g' % tuple(x) for x in p]
    lines += ['vt %.7g %.7g' % tuple(x) for x in t]
    lines += ['vn %.7g %.7g %.7g' % tuple(x) for x in n]
    lines += ['f %d/%d/%d %d/%d/%d %d/%d/%d' % (3 * i + 1, 3 * i + 1, 3 * i + 1, 3 * i + 2, 3 * i + 2, 3 * i + 2, 3 * i + 3, 3 * i + 3, 3 * i + 3) for i in range(p.shape[0] // 3)]
    return '\n'.join(lines) + '\n'


def make_assets(directory='assets'):
    os.makedirs(directory, exist_ok=True)
    out = {}
    for fname, sc in (('monkey_cornell.gltf', scenes.cornell_monkey()), ('cornell.gltf', scenes.cornell_boxes())):
        doc = gltf_document(sc['vertices'], sc['mtlids'], sc['materials'], fname[:-5])
        out[fname] = json.dumps(doc, separators=(',', ':')) + '\n'
    out['uvsphere.obj'] = obj_text(*scenes.uvsphere(32, 16, 1.0))
    for fname, text in out.items():
        with open(os.path.join(directory, fname), 'w') as fh:
            fh.write(text)
    return sorted(out)


if __name__ == '__main__':
    print(make_assets(sys.argv[1] if len(sys.argv) > 1 else 'assets'))
