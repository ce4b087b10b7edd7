"""Brute-force path tracer, no NEE / MIS (reference: ptina/engine/brute.py:17-75)."""
from .. import _native
from .path import PathEngine
from ..common import Singleton
from ..sampling.sobol import SobolSampler


class BruteEngine(metaclass=Singleton):
    ENGINE = _native.ENGINE_BRUTE

    def __init__(self):
        SobolSampler()

    render = PathEngine.render
    render_range = PathEngine.render_range
    render_tile = PathEngine.render_tile
    render_final = PathEngine.render_final
