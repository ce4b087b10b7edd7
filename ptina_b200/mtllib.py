"""Material table (reference: ptina/mtllib.py:9-95): 12 (factor vec4, texture id) pairs per material, 64 materials."""
import numpy as np

from .common import Singleton
from . import _native

SLOTS = ('basecolor', 'metallic', 'roughness', 'specular', 'specularTint', 'subsurface', 'sheen', 'sheenTint',
         'clearcoat', 'clearcoatGloss', 'transmission', 'ior')


def _expand_factor(fac):
    """ParameterPair.load's normalisation (mtllib.py:15-28): None -> 1, scalar -> x4, rgb -> rgb1."""
    if fac is None:
        fac = 1.0
    if isinstance(fac, np.ndarray):
        fac = fac.tolist() if fac.ndim else float(fac)
    if not isinstance(fac, (tuple, list)):
        return [float(fac)] * 4
    fac = [float(x) for x in fac]
    return fac + [1.0] if len(fac) == 3 else fac


def pack_materials(materials, fac=None, tex=None):
    """MaterialPool.load's table fill (mtllib.py:58-77) as a host function: each material's list is zipped with the 12 slots;
    slots a short list does not mention keep what the table held (zero-initialised: factor 0, texture id 0)."""
    materials = list(materials)
    if fac is None:
        fac, tex = np.zeros((len(materials), 12, 4), np.float32), np.zeros((len(materials), 12), np.int32)
    for i, material in enumerate(materials):
        for slot, (f, t) in zip(range(12), material):
            fac[i, slot] = _expand_factor(f)
            tex[i, slot] = t
    return fac, tex


class MaterialPool(metaclass=Singleton):
    def __init__(self, count=2**6):
        self.capacity = count
        self.count = 0
        # zero-initialised like Taichi fields: slots a short material list does not mention stay (0, tex 0)
        self.fac = np.zeros((count, 12, 4), np.float32)
        self.tex = np.zeros((count, 12), np.int32)

    def load(self, materials):
        materials = list(materials)
        assert len(materials) <= self.capacity, 'too many materials'
        pack_materials(materials, self.fac, self.tex)
        self.count = len(materials)
        _native.context().load_materials(self.fac, self.tex)
