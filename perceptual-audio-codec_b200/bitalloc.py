"""
bitalloc.py -- BitAlloc with the reference's signature and return value (codec/bitalloc.py:129-184), evaluated by the
warp-per-problem water-filling kernel (pac_bitalloc), and the three allocators HEAD's codec does not call
(BitAllocUniform / BitAllocConstSNR / BitAllocConstMNR, :22-125) on pac_bitalloc_alt.  The band layout (nLines) must be the engine's own.
"""
import numpy as np

import _pacb200


def _engine_for(nBands, nLines, who):
    e = _pacb200.engine()
    if nBands != e.nBands or list(np.asarray(nLines)) != list(e.nLines):
        e = None
        for cand in list(_pacb200._engines.values()):
            if nBands == cand.nBands and list(np.asarray(nLines)) == list(cand.nLines):
                e = cand
        if e is None:
            raise ValueError("%s: nLines is not the band layout of any open engine (psychoac.ScaleFactorBands)" % who)
    return e


def BitAllocUniform(bitBudget, maxMantBits, nBands, nLines, SMR=None):
    """bitalloc.py:22-57: equal bits per line, leftovers handed out band by band from band 0."""
    return _engine_for(nBands, nLines, "BitAllocUniform").bitalloc_alt("uniform", float(bitBudget), int(maxMantBits))[0].astype(int)


def BitAllocConstSNR(bitBudget, maxMantBits, nBands, nLines, peakSPL):
    """bitalloc.py:60-91: water-filling on a flat noise floor that starts at peakSPL in every band.  Like the reference it
    needs a budget the greedy loop can spend exactly; where the reference would spin for ever this raises."""
    e = _engine_for(nBands, nLines, "BitAllocConstSNR")
    level = np.broadcast_to(np.asarray(peakSPL, dtype=np.float64), (nBands,))      # peakSPL * ones(nBands), :70
    return e.bitalloc_alt("constsnr", float(bitBudget), int(maxMantBits), level)[0].astype(int)


def BitAllocConstMNR(bitBudget, maxMantBits, nBands, nLines, SMR):
    """bitalloc.py:94-125: water-filling on the SMRs (6 dB per bit), same termination caveat as BitAllocConstSNR."""
    e = _engine_for(nBands, nLines, "BitAllocConstMNR")
    return e.bitalloc_alt("constmnr", float(bitBudget), int(maxMantBits), np.asarray(SMR, dtype=np.float64))[0].astype(int)


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
