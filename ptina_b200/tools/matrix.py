"""4x4 float64 matrix helpers for building the `pers` (= proj @ view) matrix handed to
`Camera().set_perspective` / `worker.set_camera`.

Same function names, argument meaning and conventions (column vectors, OpenGL clip space, camera
looking down -Z) as the reference's host utilities ptina/tools/matrix.py:8-101, so driver scripts
written against PTina keep working.  Host-only NumPy; nothing here touches the GPU.
"""
import numpy as np


def identity():
    return np.identity(4)


def affine(lin, pos):
    """[lin | pos; 0 0 0 1] from a 3x3 linear part and a translation (reference: matrix.py:12-16)."""
    m = np.identity(4)
    m[:3, :3] = np.asarray(lin, dtype=float)
    m[:3, 3] = np.asarray(pos, dtype=float)
    return m


def lookat(pos=(0, 0, 0), back=(0, 0, 3), up=(0, 1, 1e-12)):
    """View matrix of a camera sitting at pos+back looking at pos (reference: matrix.py:19-31)."""
    pos, back, up = (np.asarray(v, dtype=float) for v in (pos, back, up))
    fwd = -back / np.linalg.norm(back)
    right = np.cross(fwd, up)
    right = right / np.linalg.norm(right)
    up = np.cross(right, fwd)
    cam2world = affine(np.stack([right, up, -fwd], axis=1), pos + back)
    return np.linalg.inv(cam2world)


def ortho(left=-1, right=1, bottom=-1, top=1, near=-100, far=100):
    m = np.identity(4)
    m[0, 0], m[0, 3] = 2 / (right - left), -(right + left) / (right - left)
    m[1, 1], m[1, 3] = 2 / (top - bottom), -(top + bottom) / (top - bottom)
    m[2, 2], m[2, 3] = -2 / (far - near), -(far + near) / (far - near)
    return m


def frustum(left=-1, right=1, bottom=-1, top=1, near=1, far=100):
    m = np.zeros((4, 4))
    m[0, 0], m[0, 2] = 2 * near / (right - left), (right + left) / (right - left)
    m[1, 1], m[1, 2] = 2 * near / (top - bottom), (top + bottom) / (top - bottom)
    m[2, 2], m[2, 3] = -(far + near) / (far - near), -2 * far * near / (far - near)
    m[3, 2] = -1
    return m


def orthogonal(size=1, aspect=1, near=-100, far=100):
    return ortho(-size * aspect, size * aspect, -size, size, near, far)


def perspective(fov=60, aspect=1, near=0.05, far=500):
    t = np.tan(np.radians(fov) / 2)
    return frustum(-near * t * aspect, near * t * aspect, -near * t, near * t, near, far)


def scale(factor):
    return affine(np.diag(np.ones(3) * np.asarray(factor, dtype=float)), np.zeros(3))


def translate(offset):
    return affine(np.identity(3), np.ones(3) * np.asarray(offset, dtype=float))


def quaternion(q):
    """Rotation from a unit quaternion (x, y, z, w) (reference: matrix.py:78-89)."""
    x, y, z, w = (float(c) for c in q)
    rot = np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (w * y + x * z)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])
    return affine(rot, np.zeros(3))


def eularXYZ(theta):
    """Rz @ Ry @ Rx (name spelt as in the reference, matrix.py:93-101)."""
    cx, cy, cz = np.cos(theta)
    sx, sy, sz = np.sin(theta)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return affine(rz @ ry @ rx, np.zeros(3))
