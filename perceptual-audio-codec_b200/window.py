"""
window.py -- SineWindow / HanningWindow / KBDWindow with the reference's calling convention
(codec/window.py:27-78), computed on the GPU through the C ABI (pac_window).

As in the reference, SineWindow and HanningWindow multiply their argument IN PLACE and return the same array
(window.py:37,51 -- SURVEY Appendix A Q1 depends on it); KBDWindow returns a windowed copy (window.py:62,76).
"""
import numpy as np

import _pacb200


def _eng():
    return _pacb200.engine()


def SineWindow(dataSampleArray):
    out = _eng().window(0, np.asarray(dataSampleArray, dtype=np.float64))
    dataSampleArray[...] = out
    return dataSampleArray


def HanningWindow(dataSampleArray):
    out = _eng().window(1, np.asarray(dataSampleArray, dtype=np.float64))
    dataSampleArray[...] = out
    return dataSampleArray


def KBDWindow(dataSampleArray, alpha=4.):
    if alpha != 4.:
        raise ValueError("KBDWindow: only the reference's default alpha=4 is built into the device tables")
    return _eng().window(2, np.array(dataSampleArray, dtype=np.float64))
