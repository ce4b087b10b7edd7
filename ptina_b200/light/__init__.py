"""Point and area lights (reference: ptina/light/__init__.py:9-49).  `add(world, color, size, type)` stores
pos = world @ (0,0,0,1) and axes = world[:3,:3]; the device code implements hit()/sample() (light/__init__.py:51-121)."""
import numpy as np

from ..common import Singleton
from .. import _native


class LightPool(metaclass=Singleton):
    TYPES = _native.LIGHT_TYPES

    def __init__(self, count=2**6):
        self.capacity = count
        self.count = 1          # the reference starts with one default POINT light (light/__init__.py:22-28)

    def clear(self):
        _native.context().clear_lights()
        self.count = 0

    def add(self, world, color, size, type):
        world = np.asarray(world, dtype=np.float64)
        origin = world @ np.array([0.0, 0.0, 0.0, 1.0])
        pos = origin[:3] / origin[3]
        _native.context().add_light(pos, world[:3, :3], np.asarray(color, dtype=np.float64), float(size), self.TYPES[type])
        self.count += 1
        return self.count - 1
