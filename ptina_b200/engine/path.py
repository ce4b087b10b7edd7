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

    def render_tile(self, i, j, samples):
        """One Sobol update, then samples 0..min(samples, 63) of every pixel of the 64x64 tile (i, j), each rotated by
        wanghash3(x, y, m) (path.py:96-118, commented out in the reference; driven by exams/benchtiles.py:25)."""
        _native.context().render_tile(self.ENGINE, i, j, samples)

    def render_final(self, nsamples):
        """Every tile in turn, 64 samples at a time (path.py:120-128, exams/benchtiles.py:31)."""
        _native.context().render_final(self.ENGINE, nsamples)

    def render_range(self, k_first, count, stride=1):
        """Samples at explicit Sobol point indices (for sharding a sample range over GPUs)."""
        _native.context().render_range(self.ENGINE, k_first, count, stride)
