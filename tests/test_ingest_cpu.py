"""Host-side steps either side of the hot path: glTF / GLB ingest and the daemon-thread worker shim."""
import base64
import json
import struct
import threading

import numpy as np

from ptina_b200.tools import matrix
from ptina_b200.tools.readgltf import readgltf
from ptina_b200.tools.mtworker import DaemonModule, OnDemandProxy


def _quad_doc(uri):
    pos = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    nrm = np.tile(np.array([[0, 0, 1]], np.float32), (4, 1))
    uv = pos[:, :2].copy()
    idx = np.array([0, 1, 2, 0, 2, 3], np.uint16)
    blob = pos.tobytes() + nrm.tobytes() + uv.tobytes() + idx.tobytes()
    doc = {
        'asset': {'version': '2.0'}, 'scene': 0, 'scenes': [{'nodes': [0]}],
        'nodes': [{'name': 'root', 'translation': [1.0, 2.0, 3.0], 'children': [1]},
                  {'name': 'quad', 'mesh': 0, 'scale': [2.0, 2.0, 2.0], 'rotation': [0.0, 0.0, 0.7071067811865476, 0.7071067811865476]}],
        'meshes': [{'primitives': [{'attributes': {'POSITION': 0, 'NORMAL': 1, 'TEXCOORD_0': 2}, 'indices': 3, 'material': 0}]}],
        'materials': [{'pbrMetallicRoughness': {'baseColorFactor': [0.8, 0.1, 0.1, 1.0], 'metallicFactor': 0.25, 'roughnessFactor': 0.5}}],
        'buffers': [{'byteLength': len(blob), **({'uri': uri(blob)} if uri else {})}],
        'bufferViews': [{'buffer': 0, 'byteOffset': 0, 'byteLength': 48}, {'buffer': 0, 'byteOffset': 48, 'byteLength': 48},
                        {'buffer': 0, 'byteOffset': 96, 'byteLength': 32}, {'buffer': 0, 'byteOffset': 128, 'byteLength': 12}],
        'accessors': [{'bufferView': 0, 'componentType': 5126, 'count': 4, 'type': 'VEC3'}, {'bufferView': 1, 'componentType': 5126, 'count': 4, 'type': 'VEC3'},
                      {'bufferView': 2, 'componentType': 5126, 'count': 4, 'type': 'VEC2'}, {'bufferView': 3, 'componentType': 5123, 'count': 6, 'type': 'SCALAR'}],
    }
    return doc, blob, pos, idx


def _check(vertices, mtlids, materials, images, pos, idx):
    assert vertices.shape == (6, 8) and vertices.dtype == np.float64 and list(mtlids) == [0, 0] and images == []
    world = matrix.translate([1, 2, 3]) @ (matrix.quaternion([0, 0, 0.7071067811865476, 0.7071067811865476]) @ matrix.scale([2, 2, 2]))
    want = (np.concatenate([pos[idx].astype(np.float64), np.ones((6, 1))], 1) @ world.T)[:, :3]
    assert np.allclose(vertices[:, :3], want, atol=1e-12)
    assert np.allclose(vertices[:, 3:6], [0, 0, 1], atol=1e-12)           # rotation about z leaves the normal alone
    assert np.allclose(vertices[:, 6:8], pos[idx][:, :2])
    (b, bt), (m, mt), (r, rt) = materials[0]
    assert b == [0.8, 0.1, 0.1, 1.0] and bt == -1 and (m, mt, r, rt) == (0.25, -1, 0.5, -1)


def test_gltf_embedded_and_external_buffers(tmp_path):
    doc, blob, pos, idx = _quad_doc(lambda b: 'data:application/octet-stream;base64,' + base64.b64encode(b).decode('ascii'))
    p = tmp_path / 'embedded.gltf'; p.write_text(json.dumps(doc))
    _check(*readgltf(str(p)), pos, idx)
    doc, blob, pos, idx = _quad_doc(lambda b: 'quad.bin')
    (tmp_path / 'quad.bin').write_bytes(blob)
    p = tmp_path / 'external.gltf'; p.write_text(json.dumps(doc))
    _check(*readgltf(str(p)), pos, idx)


def test_glb_container(tmp_path):
    doc, blob, pos, idx = _quad_doc(None)
    js = json.dumps(doc).encode('utf-8'); js += b' ' * (-len(js) % 4)
    bn = blob + b'\0' * (-len(blob) % 4)
    total = 12 + 8 + len(js) + 8 + len(bn)
    glb = b'glTF' + struct.pack('<II', 2, total) + struct.pack('<II', len(js), 0x4E4F534A) + js + struct.pack('<II', len(bn), 0x004E4942) + bn
    p = tmp_path / 'quad.glb'; p.write_bytes(glb)
    _check(*readgltf(str(p)), pos, idx)


def test_gltf_feeds_the_material_and_model_tables(tmp_path):
    """The reader's output is what MaterialPool.load / ModelPool.load take: 3-slot materials leave the other nine Disney
    slots at the zero defaults (mtllib.py:58-77)."""
    from ptina_b200.mtllib import pack_materials
    doc, blob, pos, idx = _quad_doc(lambda b: 'data:application/octet-stream;base64,' + base64.b64encode(b).decode('ascii'))
    p = tmp_path / 'm.gltf'; p.write_text(json.dumps(doc))
    vertices, mtlids, materials, images = readgltf(str(p))
    fac, tex = pack_materials(materials)
    assert fac.shape == (1, 12, 4) and tex.shape == (1, 12)
    assert np.allclose(fac[0, 0], [0.8, 0.1, 0.1, 1.0]) and np.allclose(fac[0, 1], 0.25) and np.allclose(fac[0, 2], 0.5)
    assert np.all(fac[0, 3:] == 0) and np.all(tex[0, :3] == -1) and np.all(tex[0, 3:] == 0)


class _FakeWorker:
    def __init__(self):
        self.threads = []
        self.value = 41

    def init(self):
        self.threads.append(threading.get_ident())
        return self

    def bump(self, by=1):
        self.threads.append(threading.get_ident())
        self.value += by
        return self.value

    def boom(self):
        raise RuntimeError('AABB step never stop! hierarchy corrupted?')


def test_daemon_module_runs_everything_on_one_thread(capsys):
    made = []
    mod = OnDemandProxy(lambda: DaemonModule(lambda: made.append(_FakeWorker()) or made[0]))
    assert made == []                                   # nothing is constructed before first use
    assert mod.value == 41                              # non-callables pass through
    mod.init()
    assert mod.bump() == 42 and mod.bump(by=8) == 50
    assert mod.boom() is None                           # swallowed and printed, like the reference (mtworker.py:31-36)
    assert 'AABB step never stop' in capsys.readouterr().out
    assert mod.bump() == 51                             # the thread survives an exception
    fake = made[0]
    assert len(set(fake.threads)) == 1 and fake.threads[0] != threading.get_ident()
    assert mod.direct_launch(threading.get_ident) == fake.threads[0]
