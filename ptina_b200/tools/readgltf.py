"""glTF 2.0 reader for the step before the hot path: `.gltf` (JSON + external / base64 buffers) and `.glb` containers ->
`(vertices, mtlids, materials, images)` in the form `worker.load_model / load_materials / load_images` take.

Same return contract as the reference's host loader (ptina/tools/readgltf.py:15-240, which sits on `gltflib`): vertices
float64 `[ntris*3, 8]` = world-space position, normalised world-space normal, TEXCOORD_0 (zeros if absent), triangles expanded
through the index accessor; `mtlids` = the primitive's material index per triangle (-1 without one); materials = 3-slot lists
`((baseColorFactor, baseColorTexture index | -1), (metallicFactor, -1), (roughnessFactor, -1))` -- the remaining nine Disney
slots stay at the table's zero defaults exactly as MaterialPool.load leaves them (mtllib.py:58-77); images = arrays
`[width, height, channels]` (x-major, like image.py expects).  Dependency-free: json + numpy (+ Pillow only if the file has
images).  A metallicRoughness texture raises, as in the reference (readgltf.py:125-126)."""
import base64
import json
import os
import struct

import numpy as np

from . import matrix

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_NCOMP = {'SCALAR': 1, 'VEC2': 2, 'VEC3': 3, 'VEC4': 4, 'MAT2': 4, 'MAT3': 9, 'MAT4': 16}


def _load_container(path):
    with open(path, 'rb') as fh:
        blob = fh.read()
    if blob[:4] == b'glTF':                                   # binary container: 12-byte header, then (length, type, data) chunks
        version, total = struct.unpack_from('<II', blob, 4)
        assert version == 2, f'unsupported GLB version {version}'
        doc, binary, pos = None, None, 12
        while pos < total:
            length, kind = struct.unpack_from('<II', blob, pos)
            data = blob[pos + 8: pos + 8 + length]
            if kind == 0x4E4F534A:
                doc = json.loads(data.decode('utf-8'))
            elif kind == 0x004E4942 and binary is None:
                binary = data
            pos += 8 + length + (-length % 4)
        assert doc is not None, 'GLB without a JSON chunk'
        return doc, binary
    return json.loads(blob.decode('utf-8')), None


def readgltf(path):
    doc, glb_bin = _load_container(path)
    base = os.path.dirname(os.path.abspath(path))

    def load_uri(uri):
        if uri.startswith('data:'):
            return base64.b64decode(uri[uri.index('base64,') + 7:].encode('ascii'))
        with open(uri if os.path.isabs(uri) else os.path.join(base, uri), 'rb') as fh:
            return fh.read()

    buffers = [glb_bin if 'uri' not in b else load_uri(b['uri']) for b in doc.get('buffers', [])]

    def view_bytes(i):
        v = doc['bufferViews'][i]
        off = v.get('byteOffset', 0)
        return buffers[v['buffer']][off: off + v['byteLength']], v.get('byteStride', 0)

    def accessor(i):
        a = doc['accessors'][i]
        dt, nc = np.dtype(_COMPONENT[a['componentType']]), _NCOMP[a['type']]
        count = a['count']
        if 'bufferView' not in a:
            return np.zeros((count, nc) if nc > 1 else count, dt)
        data, stride = view_bytes(a['bufferView'])
        off = a.get('byteOffset', 0)
        item = dt.itemsize * nc
        if stride and stride != item:                         # interleaved attributes
            raw = np.frombuffer(data, np.uint8, offset=off)
            rows = np.lib.stride_tricks.as_strided(raw, (count, item), (stride, 1))
            arr = np.ascontiguousarray(rows).view(dt).reshape(count, nc)
        else:
            arr = np.frombuffer(data, dt, count * nc, off).reshape(count, nc)
        return arr[:, 0] if nc == 1 else arr

    images = []
    for im in doc.get('images', []):
        from io import BytesIO
        from PIL import Image
        raw = load_uri(im['uri']) if 'uri' in im else view_bytes(im['bufferView'])[0]
        with BytesIO(raw) as fh:
            images.append(np.swapaxes(np.array(Image.open(fh)), 0, 1))

    materials = []
    for m in doc.get('materials', []):
        pbr = m.get('pbrMetallicRoughness', {})
        assert 'metallicRoughnessTexture' not in pbr, 'metallicRoughness texture not supported'
        bt = pbr['baseColorTexture']['index'] if 'baseColorTexture' in pbr else -1
        if bt != -1 and 'textures' in doc:                    # texture -> image source
            bt = doc['textures'][bt].get('source', bt)
        materials.append(((list(pbr.get('baseColorFactor', [1.0, 1.0, 1.0, 1.0])), bt),
                          (pbr.get('metallicFactor', 1.0), -1), (pbr.get('roughnessFactor', 1.0), -1)))

    def local_matrix(node):
        if 'matrix' in node:
            return np.array(node['matrix'], np.float64).reshape(4, 4).T        # glTF stores column-major
        mat = matrix.identity()
        if 'scale' in node:
            mat = matrix.scale(node['scale']) @ mat
        if 'rotation' in node:
            mat = matrix.quaternion(node['rotation']) @ mat
        if 'translation' in node:
            mat = matrix.translate(node['translation']) @ mat
        return mat

    arrays, mtlids = [], []

    def add_primitive(prim, world):
        if prim.get('mode', 4) != 4:
            return                                           # triangles only
        att = prim['attributes']
        p = accessor(att['POSITION']).astype(np.float64)
        assert 'NORMAL' in att, 'primitive without NORMAL'
        n = accessor(att['NORMAL']).astype(np.float64)
        t = accessor(att['TEXCOORD_0']).astype(np.float64) if 'TEXCOORD_0' in att else np.zeros((p.shape[0], 2))
        f = accessor(prim['indices']).astype(np.int64) if 'indices' in prim else np.arange(p.shape[0])
        p, n, t = p[f], n[f], t[f]
        w = np.asarray(world, np.float64).T
        ph = np.concatenate([p, np.ones((p.shape[0], 1))], 1) @ w
        p = ph[:, :3] / ph[:, 3:4]
        n = (np.concatenate([n, np.zeros((n.shape[0], 1))], 1) @ w)[:, :3]
        n = n / np.linalg.norm(n, axis=1, keepdims=True)
        a = np.concatenate([p, n, t], 1)
        assert a.shape[0] % 3 == 0
        arrays.append(a)
        mtlids.append(np.full(a.shape[0] // 3, prim.get('material', -1)))

    def walk(idx, world):
        node = doc['nodes'][idx]
        world = world @ local_matrix(node)
        if 'mesh' in node:
            for prim in doc['meshes'][node['mesh']]['primitives']:
                add_primitive(prim, world)
        for child in node.get('children', []):
            walk(child, world)

    scene = doc['scenes'][doc.get('scene', 0)]
    for root in scene['nodes']:
        walk(root, matrix.identity())
    assert arrays, 'no triangles in the glTF scene'
    return np.concatenate(arrays, 0), np.concatenate(mtlids, 0), materials, images
