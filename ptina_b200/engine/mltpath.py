"""Metropolis light transport over 2^18 chains x 32 dims (reference: ptina/engine/mltpath.py:9-87).

`LSP[None]` / `Sigma[None]` keep the reference's field-style access.  ti.random() is replaced by a counter-based
Philox stream keyed (seed, chain, iteration): statistical parity only.  `chains=(first, count)` restricts this process to
a slice of the chains (multi-GPU sharding)."""
from ..common import Singleton
from .. import _native
from ..sampling.sobol import SobolSampler


class _Scalar:
    def __init__(self, owner, name, value):
        self.owner, self.name, self.value = owner, name, value

    def __getitem__(self, key):
        return self.value

    def __setitem__(self, key, value):
        self.value = float(value)
        self.owner._push_params()


class MLTPathEngine(metaclass=Singleton):
    ENGINE = _native.ENGINE_MLT

    def __init__(self, nchains=2**18, seed=0, chains=None):
        SobolSampler()
        self.nchains, self.ndims, self.seed = nchains, 32, seed
        self.chains = chains or (0, nchains)
        self.LSP = _Scalar(self, 'LSP', 0.25)
        self.Sigma = _Scalar(self, 'Sigma', 0.01)
        self._push_params()
        self.reset()

    def _push_params(self):
        if hasattr(self, 'Sigma'):
            _native.context().mlt_set_param(self.LSP.value, self.Sigma.value)

    def reset(self):
        _native.context().mlt_reset(self.seed, self.chains[0], self.chains[1])

    def render(self, nsamples=1):
        _native.context().render(self.ENGINE, nsamples)
