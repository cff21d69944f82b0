"""
psychoac.py -- the reference's psychoacoustic entry points (codec/psychoac.py) on top of the fused SMR kernel.

  CalcSMRs / getMaskedThreshold (psychoac.py:215-318)         -> pac_calc_smrs   (mono path)
  getStereoMaskThreshold / calcBTHR (psychoac.py:409-682)     -> pac_analysis    (stereo path, what codec.Encode uses)
  ScaleFactorBands / AssignMDCTLinesFromFreqLimits (:124-213) -> host-side static layout (same arithmetic)
  SPL / Intensity / Thresh / Bark (:15-64)                    -> one-line scalar formulas kept as numpy helpers

The stereo functions take the RAW (un-windowed) time block: the reference's callers hand them buffers that have
already been sine-windowed in place; the kernel reproduces that chain itself.
"""
import numpy as np

import _pacb200

cbFreqLimits = (100.0, 200.0, 300.0, 400.0, 510.0, 630.0, 770.0, 920.0, 1080.0, 1270.0, 1480.0, 1720.0, 2000.0, 2320.0,
                2700.0, 3150.0, 3700.0, 4400.0, 5300.0, 6400.0, 7700.0, 9500.0, 12000.0, 15500.0, 24000.0)   # psychoac.py:122


def Intensity(spl):                     # psychoac.py:37-42
    return 10 ** ((spl - 96) / 10)


def SPL(intensity):                     # psychoac.py:15-35 (the in-place clamping of array arguments included)
    minval = Intensity(-30)
    if hasattr(intensity, "__len__"):
        intensity[intensity < minval] = minval
    elif intensity < minval:
        intensity = minval
    spl = 96 + 10 * np.log10(intensity)
    if hasattr(spl, "__len__"):
        spl[spl < -30.] = -30.
    elif spl < -30.:
        spl = -30.
    return spl


def Thresh(f):                          # psychoac.py:44-54
    khz = np.divide(np.clip(f, 10, np.inf), 1000.0)
    return 3.64 * (khz ** -0.8) - 6.5 * np.exp(-0.6 * ((khz - 3.3) ** 2)) + 0.001 * (khz ** 4)


def Bark(f):                            # psychoac.py:56-64
    khz = np.divide(f, 1000.0)
    return 13.0 * np.arctan(khz * 0.76) + 3.5 * np.arctan((khz / 7.5) ** 2)


def AssignMDCTLinesFromFreqLimits(nMDCTLines, sampleRate, flimit=cbFreqLimits):     # psychoac.py:124-156
    half = sampleRate // 2 if isinstance(sampleRate, (int, np.integer)) else sampleRate / 2     # Python-2 `/` on ints
    mdct_lines = (np.arange(nMDCTLines) + 0.5) / nMDCTLines * half
    lower = 0
    assignments = []
    for limit in flimit:
        upper = sampleRate / 2.0 if limit >= (sampleRate / 2.0) else limit
        inband = mdct_lines[mdct_lines <= upper]
        assignments.append(int(np.count_nonzero(inband > lower)))
        lower = upper
    return assignments


class ScaleFactorBands:                 # psychoac.py:193-213
    def __init__(self, nLines):
        self.nBands = len(nLines)
        self.nLines = np.array(nLines)
        self.lowerLine = np.append(0, np.cumsum(nLines)[:-1])
        self.upperLine = np.add(self.nLines, np.subtract(self.lowerLine, 1))


def _engine(N, sampleRate):
    return _pacb200.engine(sampleRate=int(sampleRate), nMDCTLines=N // 2)


def CalcSMRs(data, MDCTdata, MDCTscale, sampleRate, sfBands):
    """psychoac.py:253-318 (mono).  `data` is Hann-windowed in place exactly as the reference does (:225)."""
    data_in = np.array(data, dtype=np.float64)
    e = _engine(len(data_in), sampleRate)
    if list(sfBands.nLines) != list(e.nLines):
        raise ValueError("CalcSMRs: sfBands is not the layout AssignMDCTLinesFromFreqLimits gives for this N / sampleRate")
    smr = e.calc_smrs(data_in, np.asarray(MDCTdata, dtype=np.float64), int(MDCTscale))[0]
    try:
        data[...] = e.window(1, data_in)          # the reference leaves its argument Hann-windowed
    except TypeError:
        pass
    return smr


def getStereoMaskThreshold(data, MDCTdata, MDCTscale, sampleRate, sfBands, LRMS, codingParams=None):
    """psychoac.py:506-682.  Returns (SMR[2][nBands], LRMSmdctLines[2][nMDCTLines]).

    Takes the raw time blocks (see the module docstring); MDCTdata / MDCTscale / LRMS are recomputed by the fused
    kernel from `data` and must agree with what the caller passed (they do when they come from codec.Encode)."""
    blk = np.stack([np.asarray(data[0], dtype=np.float64), np.asarray(data[1], dtype=np.float64)])
    e = _engine(blk.shape[1], sampleRate)
    r = e.analysis(blk[None])
    mask = sum((1 << b) for b, v in enumerate(LRMS) if v)
    if int(r["lrms"][0]) != mask:
        raise ValueError("getStereoMaskThreshold: LRMS differs from the decision the kernel derives from `data`")
    return r["smr"][0], r["lines"][0]


def getMaskedThreshold(data, MDCTdata, MDCTscale, sampleRate, sfBands):
    """psychoac.py:215-251: masked threshold (dB) at the MDCT lines; `data` is left Hann-windowed like the reference."""
    return calcBTHR(data, MDCTdata, MDCTscale, sampleRate, sfBands, False)


def calcBTHR(data, MDCTdata, MDCTscale, sampleRate, sfBands, noDrop):
    """psychoac.py:409-456: Hann window (IN PLACE, :428) -> FFT -> peaks -> maskers -> spreading -> + threshold in quiet."""
    x = np.array(data, dtype=np.float64)
    e = _engine(len(x), sampleRate)
    thr = e.masked_threshold(x, noDrop)[0]
    try:
        data[...] = e.window(1, x)
    except TypeError:
        pass
    return thr
