"""Film accumulator (reference: ptina/filmtable.py:8-79): 3 passes x 2^21 float4 running sums (rgb, sample count),
pixel (x, y) at index x*ny + y."""
import numpy as np

from .common import Singleton
from . import _native


class FilmTable(metaclass=Singleton):
    def __init__(self, size=2**21, count=3):
        self.size, self.count = size, count

    @property
    def nx(self):
        return _native.context().get_size()[0]

    @property
    def ny(self):
        return _native.context().get_size()[1]

    def set_size(self, nx, ny):
        _native.context().set_size(nx, ny)

    def clear(self, id=0):
        _native.context().clear()           # zeroes ALL passes, like filmtable.py:44-45

    def get_image(self, id=0):
        """New float32 [nx, ny, 4] array: rgb / w, a = 1; (0.9, 0.4, 0.9, 0) where no sample landed (filmtable.py:47-63)."""
        return _native.context().get_image(id)

    def fast_export_image(self, out, id=0):
        """Fill the caller's float32 [ny*nx*3] buffer, row-major rgb (filmtable.py:65-79)."""
        _native.context().fast_export_image(out, id)

    def to_torch(self, id=0):
        return _native.context().film_tensor(id)
