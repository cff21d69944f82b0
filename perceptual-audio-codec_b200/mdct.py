"""
mdct.py -- MDCT / IMDCT with the reference's signature (codec/mdct.py:49-88) on the GPU (pac_mdct / pac_imdct):
fold + N/4-point complex FFT in shared memory instead of the reference's N-point complex FFT.
Only the symmetric case a == b that the codec uses is supported.
"""
import numpy as np

import _pacb200


def MDCT(data, a, b, isInverse=False):
    if a != b:
        raise ValueError("MDCT: only a == b (the codec's case, codec.py:241) is implemented on the device")
    data = np.asarray(data, dtype=np.float64)
    e = _pacb200.engine()
    if not isInverse:
        if data.shape[-1] != a + b:
            raise ValueError("MDCT: expected %d samples" % (a + b))
        return e.mdct(data)
    if data.shape[-1] != (a + b) // 2:
        raise ValueError("IMDCT: expected %d lines" % ((a + b) // 2))
    return e.imdct(data)


def IMDCT(data, a, b):
    return MDCT(data, a, b, True)


def MDCTslow(data, a, b, isInverse=False):
    """mdct.py:10-43 computes the same transform in O(N^2); the device path is exact to rounding."""
    return MDCT(data, a, b, isInverse)
