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
        for i, material in enumerate(materials):
            for slot, (fac, tex) in zip(range(12), material):
                self.fac[i, slot] = _expand_factor(fac)
                self.tex[i, slot] = tex
        self.count = len(materials)
        _native.context().load_materials(self.fac, self.tex)
