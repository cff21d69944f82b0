"""
bitalloc.py -- BitAlloc with the reference's signature and return value (codec/bitalloc.py:129-184), evaluated by the
warp-per-problem water-filling kernel (pac_bitalloc).  The band layout (nLines) must be the engine's own.
"""
import numpy as np

import _pacb200


def BitAlloc(bitBudget, extraBits, maxMantBits, nBands, nLines, SMR, LRMS):
    e = _pacb200.engine()
    if nBands != e.nBands or list(np.asarray(nLines)) != list(e.nLines):
        e = None
        for cand in list(_pacb200._engines.values()):
            if nBands == cand.nBands and list(np.asarray(nLines)) == list(cand.nLines):
                e = cand
        if e is None:
            raise ValueError("BitAlloc: nLines is not the band layout of any open engine (psychoac.ScaleFactorBands)")
    mask = 0
    for b, v in enumerate(LRMS):
        if v:
            mask |= 1 << b
    bits, diff = e.bitalloc(float(bitBudget), int(extraBits), int(maxMantBits), np.asarray(SMR, dtype=np.float64), mask)
    return bits[0].astype(int), int(diff[0])
