"""Flat function API -- the drop-in boundary scripts and the Blender add-on call (through tools/mtworker.py).

Same names, argument meaning and error behaviour as the reference's ptina/worker.py:11-87.  Most entries are plain forwards to
one method of one singleton; they are generated from the table below (name -> singleton, method, reference line) so that the
mapping to the reference is explicit in one place.  Extras over the reference: `init(**caps)` takes init_things' capacities,
`render / render_preview(nsamples=)` batch several samples into one wavefront submission, `synchronize()` really waits for
the device.
"""
import functools

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

# name: (singleton class, method, parameter names with defaults, returns a value?, reference worker.py line)
_FORWARDS = {
    'set_size':          (FilmTable,    'set_size',          ('nx', 'ny'),                        False, 29),
    'get_image':         (FilmTable,    'get_image',         (('id', 0),),                        True,  50),
    'fast_export_image': (FilmTable,    'fast_export_image', ('pixels', ('id', 0)),               False, 54),
    'clear_lights':      (LightPool,    'clear',             (),                                  False, 58),
    'set_world_light':   (WorldLight,   'set',               ('fac', 'tex'),                      False, 62),
    'add_light':         (LightPool,    'add',               ('world', 'color', 'size', 'type'),  False, 66),
    'load_model':        (ModelPool,    'load',              ('vertices', 'mtlids'),              False, 70),
    'load_images':       (ImagePool,    'load',              ('images',),                         False, 74),
    'load_materials':    (MaterialPool, 'load',              ('materials',),                      False, 78),
    'build_tree':        (BVHTree,      'build',             (),                                  False, 82),
    'set_camera':        (Camera,       'set_perspective',   ('pers',),                           False, 86),
}


def _forward(name, cls, method, params, returns, line):
    names = [p if isinstance(p, str) else p[0] for p in params]
    defaults = tuple(p[1] for p in params if not isinstance(p, str))

    def call(*args, **kwargs):
        if len(args) > len(names):
            raise TypeError(f'{name}() takes {len(names)} positional arguments but {len(args)} were given')
        bound = dict(zip(names, args))
        for key, value in kwargs.items():
            if key not in names or key in bound:
                raise TypeError(f'{name}() got an unexpected or repeated argument {key!r}')
            bound[key] = value
        for key, value in zip(names[len(names) - len(defaults):], defaults):
            bound.setdefault(key, value)
        missing = [k for k in names if k not in bound]
        if missing:
            raise TypeError(f'{name}() missing required argument {missing[0]!r}')
        result = getattr(cls(), method)(*[bound[k] for k in names])
        return result if returns else None

    call.__name__ = call.__qualname__ = name
    call.__doc__ = f'{cls.__name__}().{method}({", ".join(names)})  -- ptina/worker.py:{line}'
    return call


for _name, _spec in _FORWARDS.items():
    globals()[_name] = _forward(_name, *_spec)
del _name, _spec


def init(**caps):
    """worker.py:11-14: create the singletons (init_things) and the two engines the module drives."""
    init_things(**caps)
    for engine in (DefaultEngine, PreviewEngine):
        engine()


def synchronize():
    """worker.py:17-18 (the reference only prints the face count there): wait until the device has finished."""
    _native.context().synchronize()


def render(aa=True, nsamples=1):
    """worker.py:21-22: one more sample per pixel from the default engine (nsamples > 1: a batch in one submission)."""
    DefaultEngine().render(nsamples)


def render_preview(aa=True, nsamples=1):
    """worker.py:25-26: albedo / normal passes (film 1 and 2)."""
    PreviewEngine().render(nsamples)


def get_size():
    """worker.py:33-34."""
    film = FilmTable()
    return film.nx, film.ny


def clear(id=0):
    """worker.py:37-40: engines with per-run state (MLT) restart, then every film pass is zeroed (filmtable.py:44-45)."""
    engine = DefaultEngine()
    reset = getattr(engine, 'reset', None)
    if reset is not None:
        reset()
    FilmTable().clear(id)


def set_mlt_param(lsp, sigma):
    """worker.py:43-47: large-step probability / small-step sigma of the Metropolis engine, ignored by engines without them."""
    engine = DefaultEngine()
    for field, value in (('LSP', lsp), ('Sigma', sigma)):
        holder = getattr(engine, field, None)
        if holder is not None:
            holder[None] = value
