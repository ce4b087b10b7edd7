"""`init_things()` -- allocates every pool with the reference's capacities (ptina/things.py:12-28)."""
from . import _native
from .camera import Camera
from .tree import BVHTree
from .image import ImagePool
from .model import ModelPool
from .light import LightPool
from .light.world import WorldLight
from .mtllib import MaterialPool
from .filmtable import FilmTable

_POOLS = (Camera, BVHTree, ImagePool, ModelPool, LightPool, WorldLight, MaterialPool, FilmTable)


def init_things(max_faces=2**21, max_texels=2**22, max_materials=2**6, max_textures=2**6, max_lights=2**6,
                max_filmsize=2**21, max_filmpasses=3, max_paths=0, device=None):
    _native.context(max_faces=max_faces, max_texels=max_texels, max_materials=max_materials, max_textures=max_textures,
                    max_lights=max_lights, max_filmsize=max_filmsize, max_filmpasses=max_filmpasses, max_paths=max_paths,
                    **({'device': device} if device is not None else {}))
    Camera()
    BVHTree(max_faces)
    ImagePool(max_texels, max_textures)
    ModelPool(max_faces)
    LightPool(max_lights)
    WorldLight()
    MaterialPool(max_materials)
    FilmTable(max_filmsize, max_filmpasses)


def shutdown():
    """Drop the native context and every singleton (tests use this to start from a clean process state)."""
    from .sampling.sobol import SobolSampler
    from . import engine
    for cls in _POOLS + (SobolSampler,) + engine.ENGINES:
        cls._forget()
    _native.shutdown()
