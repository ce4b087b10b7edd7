"""Texture pool (reference: ptina/image.py:9-95): up to 64 images in one float4 arena, x-major texels, ids assigned in
load order; uint8 input is divided by 255, grey -> rgb, rgb -> rgba(1) (image.py:69-89)."""
import numpy as np

from .common import Singleton
from .allocator import MemoryAllocator, IdAllocator
from . import _native


class ImagePool(metaclass=Singleton):
    def __init__(self, size=2**22, count=2**6):
        self.mman = MemoryAllocator(size)
        self.idman = IdAllocator(count)
        self.nx = np.zeros(count, np.int32)
        self.ny = np.zeros(count, np.int32)
        self.base = np.zeros(count, np.int32)
        self._pending = []

    @staticmethod
    def _to_rgba(arr):
        if isinstance(arr, str):
            from PIL import Image as PILImage
            arr = np.asarray(PILImage.open(arr)).swapaxes(0, 1)[:, ::-1]     # ti.imread orientation: x right, y up
        arr = np.asarray(arr)
        if arr.dtype == np.uint8:
            arr = arr.astype(np.float32) / 255
        if arr.ndim == 2:
            arr = arr[:, :, None]
        if arr.shape[2] == 1:
            arr = np.repeat(arr, 3, axis=2)
        if arr.shape[2] == 3:
            arr = np.concatenate([arr, np.ones(arr.shape[:2] + (1,), arr.dtype)], axis=2)
        return np.ascontiguousarray(arr, dtype=np.float32)

    def load_one(self, arr):
        rgba = self._to_rgba(arr)
        nx, ny = rgba.shape[:2]
        id = self.idman.malloc()
        base = self.mman.malloc(nx * ny)
        self.nx[id], self.ny[id], self.base[id] = nx, ny, base
        self._pending.append((base, rgba.reshape(nx * ny, 4)))
        return id

    def load(self, images):
        self.mman.reset()
        self.idman.reset()
        self._pending = []
        for arr in images:
            self.load_one(arr)
        self.flush()

    def flush(self):
        n = self.idman.water
        ntex = max((b + c.shape[0] for b, c in self._pending), default=0)
        arena = np.zeros((max(ntex, 1), 4), np.float32)
        for b, c in self._pending:
            arena[b:b + c.shape[0]] = c
        _native.context().load_images(arena[:ntex], self.nx[:n], self.ny[:n], self.base[:n])
