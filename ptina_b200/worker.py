"""Flat function API (reference: ptina/worker.py:11-87) -- the drop-in boundary the Blender add-on and scripts call."""
from . import _native
from .things import init_things
from .engine import PathEngine as DefaultEngine, PreviewEngine
from .tree import BVHTree
from .image import ImagePool
from .model import ModelPool
from .light import LightPool
from .light.world import WorldLight
from .mtllib import MaterialPool
from .filmtable import FilmTable
from .camera import Camera


def init(**caps):
    init_things(**caps)
    DefaultEngine()
    PreviewEngine()


def synchronize():
    _native.context().synchronize()


def render(aa=True, nsamples=1):
    DefaultEngine().render(nsamples)


def render_preview(aa=True, nsamples=1):
    PreviewEngine().render(nsamples)


def set_size(nx, ny):
    FilmTable().set_size(nx, ny)


def get_size():
    return FilmTable().nx, FilmTable().ny


def clear(id=0):
    if hasattr(DefaultEngine(), 'reset'):
        DefaultEngine().reset()
    FilmTable().clear(id)


def set_mlt_param(lsp, sigma):
    if hasattr(DefaultEngine(), 'LSP'):
        DefaultEngine().LSP[None] = lsp
    if hasattr(DefaultEngine(), 'Sigma'):
        DefaultEngine().Sigma[None] = sigma


def get_image(id=0):
    return FilmTable().get_image(id)


def fast_export_image(pixels, id=0):
    FilmTable().fast_export_image(pixels, id)


def clear_lights():
    LightPool().clear()


def set_world_light(fac, tex):
    WorldLight().set(fac, tex)


def add_light(world, color, size, type):
    LightPool().add(world, color, size, type)


def load_model(vertices, mtlids):
    ModelPool().load(vertices, mtlids)


def load_images(images):
    ImagePool().load(images)


def load_materials(materials):
    MaterialPool().load(materials)


def build_tree():
    BVHTree().build()


def set_camera(pers):
    Camera().set_perspective(pers)
