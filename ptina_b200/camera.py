"""Camera (reference: ptina/camera.py:8-39).  `set_perspective(pers)` takes the 4x4 float64 world->clip matrix; its
float64 inverse is computed on the host exactly as the reference does and both are stored as float32."""
import numpy as np

from .common import Singleton
from . import _native
from .tools import matrix


class Camera(metaclass=Singleton):
    def __init__(self):
        self.set_perspective(matrix.ortho() @ matrix.lookat())      # camera.py:14-17 default

    def set_perspective(self, pers):
        pers = np.asarray(pers, dtype=np.float64)
        self.W2V = pers.astype(np.float32)
        self.V2W = np.linalg.inv(pers).astype(np.float32)
        _native.context().set_camera(self.V2W, self.W2V)
