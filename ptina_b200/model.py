"""Triangle-mesh storage (reference: ptina/model.py:8-101).  `ModelPool().load(arr, mtlids)` takes the same inputs:
vertices float32/64 [nfaces*3, 8] = px py pz nx ny nz u v (or an OBJ dict / path), mtlids int32 [nfaces]
(-1 = default material).  The arrays are copied into device memory owned by the native context."""
import numpy as np

from .common import Singleton
from . import _native


class ModelPool(metaclass=Singleton):
    def __init__(self, size=2**21):
        self.size = size

    @property
    def nfaces(self):
        return _native.context().nfaces

    def load(self, arr, mtlids=None):
        if isinstance(arr, str):
            from .tools.readobj import readobj
            arr = readobj(arr)
        if isinstance(arr, dict):
            f = arr['f']
            arr = np.concatenate([arr['v'][f[:, :, 0]].reshape(-1, 3), arr['vn'][f[:, :, 2]].reshape(-1, 3),
                                  arr['vt'][f[:, :, 1]].reshape(-1, 2)], axis=1)
        on_device = _native._is_cuda_tensor(arr)
        if not on_device:
            arr = np.asarray(arr)
            if arr.dtype != np.float32:
                arr = arr.astype(np.float32)
            arr = np.ascontiguousarray(arr)
        assert arr.shape[0] % 3 == 0
        if mtlids is None:
            mtlids = -np.ones(arr.shape[0] // 3, dtype=np.int32)
            if on_device:
                import torch
                mtlids = torch.from_numpy(mtlids).to(arr.device)
        else:
            assert mtlids.shape[0] == arr.shape[0] // 3
            if not on_device:
                mtlids = np.ascontiguousarray(mtlids, dtype=np.int32)
        assert mtlids.shape[0] < self.size, 'too many faces'
        _native.context().load_model(arr, mtlids)

    from_numpy = load
