"""Albedo / normal AOV tracer (reference: ptina/engine/preview.py:11-41): film pass 1 += albedo, pass 2 += normal."""
from .. import _native
from .path import PathEngine
from ..common import Singleton
from ..sampling.sobol import SobolSampler


class PreviewEngine(metaclass=Singleton):
    ENGINE = _native.ENGINE_PREVIEW

    def __init__(self):
        SobolSampler()

    render = PathEngine.render
    render_range = PathEngine.render_range
