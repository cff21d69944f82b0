"""
quantize.py -- the reference's quantiser entry points (codec/quantize.py) on the GPU (pac_* in include/pac_b200.h).
Integer results are bit-exact; argument order and return types follow the reference (uint64 code vectors).
"""
import numpy as np

import _pacb200


def _e():
    return _pacb200.engine()


def vQuantizeUniform(aNumVec, nBits):                               # quantize.py:91-117
    return _e().vquantize_uniform(np.asarray(aNumVec, dtype=np.float64), int(nBits))


def vDequantizeUniform(aQuantizedNumVec, nBits):                    # quantize.py:120-145
    return _e().vdequantize_uniform(np.asarray(aQuantizedNumVec).astype(np.uint64), int(nBits))


def QuantizeUniform(aNum, nBits):                                   # quantize.py:40-64
    if nBits <= 0:
        return 0
    return int(vQuantizeUniform(np.array([aNum], dtype=np.float64), nBits)[0])


def DequantizeUniform(aQuantizedNum, nBits):                        # quantize.py:67-88
    if nBits <= 0:
        return 0
    return float(vDequantizeUniform(np.array([aQuantizedNum], dtype=np.uint64), nBits)[0])


def ScaleFactor(aNum, nScaleBits=3, nMantBits=5):                   # quantize.py:148-177
    if nMantBits <= 0:
        return 0
    return int(_e().scale_factor(float(aNum), int(nScaleBits), int(nMantBits))[0])


def vMantissa(aNumVec, scale, nScaleBits=3, nMantBits=5):           # quantize.py:315-342
    return _e().vmantissa(np.asarray(aNumVec, dtype=np.float64), int(scale), int(nScaleBits), int(nMantBits))


def vDequantize(scale, mantissaVec, nScaleBits=3, nMantBits=5):     # quantize.py:345-376
    return _e().vdequantize(int(scale), np.asarray(mantissaVec).astype(np.int64), int(nScaleBits), int(nMantBits))


def Mantissa(aNum, scale, nScaleBits=3, nMantBits=5):               # quantize.py:249-277
    if nMantBits <= 0:
        return 0.0
    return int(vMantissa(np.array([aNum], dtype=np.float64), scale, nScaleBits, nMantBits)[0])


def Dequantize(scale, mantissa, nScaleBits=3, nMantBits=5):         # quantize.py:280-312
    if nMantBits <= 0:
        return 0
    return float(vDequantize(scale, np.array([mantissa], dtype=np.int64), nScaleBits, nMantBits)[0])
