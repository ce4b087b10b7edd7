"""Environment light (reference: ptina/light/world.py:9-29): colour factor x optional equirect texture."""
from ..common import Singleton
from .. import _native


class WorldLight(metaclass=Singleton):
    def __init__(self):
        self.fac, self.tex = [0.1] * 4, 0     # world.py:12-16 (the tex field is zero-initialised, not -1)

    def set(self, fac, tex):
        self.fac, self.tex = list(fac), int(tex)
        _native.context().set_world_light(self.fac, self.tex)
