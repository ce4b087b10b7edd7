"""Hashing helpers of the sampler (reference: ptina/sampling/__init__.py:8-23), host-side mirrors of the device code."""
import numpy as np


def wanghash(x):
    """u32 Wang hash exactly as the reference writes it (note `v ^= v << 4` is a LEFT shift there); returns i32."""
    v = np.uint32(np.int64(x) & 0xFFFFFFFF)
    with np.errstate(over='ignore'):
        v = (v ^ np.uint32(61)) ^ (v >> np.uint32(16))
        v = v * np.uint32(9)
        v = v ^ (v << np.uint32(4))
        v = v * np.uint32(0x27d4eb2d)
        v = v ^ (v >> np.uint32(15))
    return int(np.int32(v))


def wanghash2(x, y):
    return wanghash(y ^ wanghash(x))
