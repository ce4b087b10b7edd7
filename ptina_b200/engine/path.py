"""Unidirectional path tracer with next-event estimation and MIS (reference: ptina/engine/path.py:10-93)."""
from ..common import Singleton
from .. import _native
from ..sampling.sobol import SobolSampler


class PathEngine(metaclass=Singleton):
    ENGINE = _native.ENGINE_PATH

    def __init__(self):
        SobolSampler()

    def render(self, nsamples=1):
        """`nsamples` x (SobolSampler().update(); _render()) (path.py:75-77)."""
        _native.context().render(self.ENGINE, nsamples)

    def render_range(self, k_first, count, stride=1):
        """Samples at explicit Sobol point indices (for sharding a sample range over GPUs)."""
        _native.context().render_range(self.ENGINE, k_first, count, stride)
