"""
ref_py3.py -- run the UNMODIFIED reference algorithm under Python 3 / NumPy 2.

TEST INFRASTRUCTURE ONLY (never imported by the product path).

The reference (`/root/reference/codec/*.py`) is Python-2.7 / NumPy-1.x code and
cannot be imported by this image's interpreter.  This module materialises a
*mechanically patched* copy into a throw-away temp directory (never into the
repo), imports it from there, and hands back the modules.  The patch list is the
one verified in SURVEY.md Appendix C: it only touches Python-2 syntax and
NumPy-2 promotion hazards, never the arithmetic.  Every substitution asserts how
many times it fired, so a silently un-applied patch fails loudly.

It is used by `oracle/make_golden.py` (here, in the build container) to
  * prove that the patched reference reproduces the committed goldens
    (`coded/*.wak`, `outputs/*.wav`) byte for byte, and
  * generate the fixtures in `tests/golden/` that pin `oracle/pac_oracle.c`.

`/root/reference` does not exist on the GPU box; nothing at test/bench time
imports this file unless the reference tree is present.
"""
import importlib
import os
import re
import shutil
import sys
import tempfile

REF_ROOT = os.environ.get("PAC_REFERENCE_ROOT", "/root/reference")
REF_CODEC = os.path.join(REF_ROOT, "codec")

MODULES = ["audiofile", "bitalloc", "bitpack", "codec", "Huffman", "mdct",
           "pacfile", "pcmfile", "psychoac", "quantize", "window"]

_BUILTINS = ("import builtins as _b\n"
             "max=_b.max; min=_b.min; abs=_b.abs; round=_b.round; all=_b.all; "
             "any=_b.any; pow=_b.pow; bool=_b.bool\n")

# (regex, replacement, expected number of substitutions)
PATCHES = {
    "bitpack": [
        (r"xrange", "range", 2),
        (r"self\.data\.tostring\(\)", "self.data.tobytes()", 1),
        (r"np\.fromstring\(data,dtype=np\.uint8\)", "np.frombuffer(data,dtype=np.uint8)", 1),
        # NumPy-2 keeps uint8 through & and << and silently wraps
        (r"dataMask &= self\.data\[self\.iByte\]", "dataMask &= int(self.data[self.iByte])", 2),
        (r"dataMask = self\.data\[self\.iByte\]", "dataMask = int(self.data[self.iByte])", 1),
        (r"def WriteBits\(self,info,nBits\):\n(\s+\"\"\".*?\"\"\"\n)",
         r"def WriteBits(self,info,nBits):\n\1        info=int(info); nBits=int(nBits)\n", 1),
        (r"def ReadBits\(self,nBits\):\n(\s+\"\"\".*?\"\"\"\n)",
         r"def ReadBits(self,nBits):\n\1        nBits=int(nBits)\n", 1),
    ],
    "mdct": [
        (r"np\.arange\(0,N/2\)", "np.arange(0,N//2)", 2),
        (r"np\.zeros\(N/2\)", "np.zeros(N//2)", 1),
        (r"fft\[0:N/2\]", "fft[0:N//2]", 1),
    ],
    "window": [
        (r"from mdct import \*\n", "from mdct import *\n" + _BUILTINS, 1),
    ],
    "psychoac": [
        (r"(?m)^import solution\.psychoac_ as sol\n", _BUILTINS, 1),
        (r"\[0:N/2\]", "[0:N//2]", 2),
        (r"p = \(1/2\)\*", "p = (1//2)*", 1),
        (r"\*\(sampleRate/N\)\)", "*(sampleRate//N))", 1),
        (r"Xwdb\[freqs\[idx\]-1\]", "Xwdb[int(freqs[idx])-1]", 1),
        (r"Xwdb\[freqs\[idx\]\]", "Xwdb[int(freqs[idx])]", 1),
        (r"Xwdb\[freqs\[idx\]\+1\]", "Xwdb[int(freqs[idx])+1]", 1),
    ],
    "quantize": [
        (r"def QuantizeUniform\(aNum,nBits\):\n(\s+\"\"\".*?\"\"\"\n)",
         r"def QuantizeUniform(aNum,nBits):\n\1    nBits=int(nBits)\n", 1),
        (r"def vQuantizeUniform\(aNumVec, nBits\):\n(\s+\"\"\".*?\"\"\"\n)",
         r"def vQuantizeUniform(aNumVec, nBits):\n\1    nBits=int(nBits)\n", 1),
        (r"def vDequantizeUniform\(aQuantizedNumVec, nBits\):\n(\s+\"\"\".*?\"\"\"\n)",
         r"def vDequantizeUniform(aQuantizedNumVec, nBits):\n\1    nBits=int(nBits)\n", 1),
        (r"def ScaleFactor\(aNum, nScaleBits=3, nMantBits=5\):\n(\s+\"\"\".*?\"\"\"\n)",
         r"def ScaleFactor(aNum, nScaleBits=3, nMantBits=5):\n\1    nScaleBits=int(nScaleBits); nMantBits=int(nMantBits)\n", 1),
        (r"def vMantissa\(aNumVec, scale, nScaleBits=3, nMantBits=5\):\n(\s+\"\"\".*?\"\"\"\n)",
         r"def vMantissa(aNumVec, scale, nScaleBits=3, nMantBits=5):\n\1    scale=int(scale); nScaleBits=int(nScaleBits); nMantBits=int(nMantBits)\n", 1),
        (r"def vDequantize\(scale, mantissaVec, nScaleBits=3, nMantBits=5\):\n(\s+\"\"\".*?\"\"\"\n)",
         r"def vDequantize(scale, mantissaVec, nScaleBits=3, nMantBits=5):\n\1    scale=int(scale); nScaleBits=int(nScaleBits); nMantBits=int(nMantBits)\n", 1),
    ],
    "bitalloc": [
        (r"from psychoac import \*\n", "from psychoac import *\n" + _BUILTINS, 1),
        (r"print '\*'", "print('*', end='')", 1),
    ],
    "Huffman": [
        (r"import cPickle as pickle", "import pickle", 1),
        (r"self\.bitDeposit/100", "self.bitDeposit//100", 1),
    ],
    "codec": [
        (r"from Huffman import \*\n", "from Huffman import *\n" + _BUILTINS + "_bsum=_b.sum\n", 1),
        (r"sum\(len\(huff\) for huff in m\)", "_bsum(len(huff) for huff in m)", 1),
    ],
    "pcmfile": [
        (r'!= "RIFF"', '!= b"RIFF"', 1),
        (r'!= "WAVE"', '!= b"WAVE"', 1),
        (r'== "fmt "', '== b"fmt "', 1),
        (r'== "data"', '== b"data"', 1),
        (r'raise "ERROR', 'raise Exception("ERROR', 2),
        (r'RIFF file!"\n', 'RIFF file!")\n', 1),
        (r'RIFF file header"\n', 'RIFF file header")\n', 1),
        (r"numSamples /= nChannels", "numSamples //= nChannels", 1),
        (r"bitsPerSample/BYTESIZE", "bitsPerSample//BYTESIZE", 9),
        (r'\*"\\0"', r'*b"\\0"', 1),
        (r"xrange", "range", 1),
        (r"dataBlock\.tostring\(\)", "dataBlock.tobytes()", 1),
        (r'"RIFF", chunkSize, "WAVE", "fmt "', 'b"RIFF", chunkSize, b"WAVE", b"fmt "', 1),
        (r'bitsPerSample, "data", dataBytes', 'bitsPerSample, b"data", dataBytes', 1),
    ],
    "pacfile": [
        (r"tag='PAC '", "tag=b'PAC '", 1),
        (r'raise "Tried to read a non-PAC file into a PACFile object"',
         'raise Exception("Tried to read a non-PAC file into a PACFile object")', 1),
        (r'raise "Only read a partial block of coded PACFile data"',
         'raise Exception("Only read a partial block of coded PACFile data")', 1),
        (r"nBytes /= BYTESIZE", "nBytes //= BYTESIZE", 1),
        (r"nBytes = nBytes/BYTESIZE \+ 1", "nBytes = nBytes//BYTESIZE + 1", 1),
        (r"dtype=np\.float\)", "dtype=float)", 2),
        # Close() relies on a module-global `huffman` set by __main__; expose a settable global
        (r"MAX16BITS = 32767\n", "MAX16BITS = 32767\nhuffman = None\n", 1),
    ],
    "audiofile": [],
}

_STUBS = {
    "matplotlib/__init__.py": "",
    "matplotlib/pyplot.py": "",
    "pylab.py": "",
}


def _patch_source(name, text):
    # drop the Python-2 `__main__` test scaffolding (print statements)
    m = re.search(r"^if __name__\s*==\s*['\"]__main__['\"]\s*:", text, flags=re.M)
    if m:
        text = text[:m.start()]
    # pcmfile has a doc-string block before the imports that is fine; BOMs are not
    text = text.lstrip("﻿")
    for pat, rep, count in PATCHES[name]:
        text, n = re.subn(pat, rep, text, flags=re.S)
        if n != count:
            raise RuntimeError("ref_py3: patch %r on %s.py fired %d times, expected %d"
                               % (pat, name, n, count))
    return text


def available():
    return os.path.isdir(REF_CODEC)


def materialize(dst=None):
    """Write the patched modules + stubs + pickles into `dst` (default: a fresh
    temp dir OUTSIDE the repo) and return the path."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_CODEC)
    if dst is None:
        dst = tempfile.mkdtemp(prefix="pac_ref_py3_")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if os.path.abspath(dst).startswith(repo + os.sep):
        raise RuntimeError("refusing to materialise reference sources inside the repo")
    os.makedirs(dst, exist_ok=True)
    for name in MODULES:
        with open(os.path.join(REF_CODEC, name + ".py"), encoding="utf-8") as f:
            text = f.read()
        with open(os.path.join(dst, name + ".py"), "w", encoding="utf-8") as f:
            f.write(_patch_source(name, text))
    for rel, body in _STUBS.items():
        p = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(body)
    for pk in ("huffmanTables.pickle", "histograms.pickle"):
        shutil.copyfile(os.path.join(REF_CODEC, pk), os.path.join(dst, pk))
    return dst


class Ref:
    """Handle on an imported, patched reference.  Use as a context manager or
    call close() to delete the temp copy."""

    def __init__(self):
        self.dir = materialize()
        self._saved_path = list(sys.path)
        self._saved_cwd = os.getcwd()
        clash = [m for m in MODULES + ["matplotlib", "pylab"] if m in sys.modules]
        self._saved_modules = {m: sys.modules.pop(m) for m in clash}
        sys.path.insert(0, self.dir)
        os.chdir(self.dir)  # Huffman() opens its pickles relative to CWD (Huffman.py:257-260)
        try:
            self.mod = {m: importlib.import_module(m) for m in MODULES}
        finally:
            os.chdir(self._saved_cwd)
        for k, v in self.mod.items():
            setattr(self, k, v)

    def new_huffman(self):
        cwd = os.getcwd()
        os.chdir(self.dir)
        try:
            return self.Huffman.Huffman()
        finally:
            os.chdir(cwd)

    def close(self):
        for m in MODULES + ["matplotlib", "matplotlib.pyplot", "pylab"]:
            sys.modules.pop(m, None)
        sys.modules.update(self._saved_modules)
        sys.path[:] = self._saved_path
        shutil.rmtree(self.dir, ignore_errors=True)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- whole-file drivers, following pacfile.py:430-499 ----
    def encode_file(self, wav_in, pac_out, target_bits=2.27, hook=None):
        """PCM wav -> .pac/.wak, exactly as the reference `__main__` Encode pass
        (pacfile.py:434-499).  Returns (bitDeposit, extraBits) at the end."""
        huffman = self.new_huffman()
        self.pacfile.huffman = huffman
        inFile = self.pcmfile.PCMFile(wav_in)
        outFile = self.pacfile.PACFile(pac_out)
        cp = inFile.OpenForReading()
        cp.nMDCTLines = 1024
        cp.nScaleBits = 4
        cp.nMantSizeBits = 4
        cp.targetBitsPerSample = target_bits
        cp.nTableIDBits = 4
        cp.nSamplesPerBlock = cp.nMDCTLines
        outFile.OpenForWriting(cp)
        iblk = 0
        while True:
            data = inFile.ReadDataBlock(cp)
            if not data:
                break
            if hook is not None:
                hook(iblk, data, cp, huffman)
            outFile.WriteDataBlock(data, cp, huffman)
            iblk += 1
        inFile.Close(cp)
        outFile.Close(cp)
        return huffman.bitDeposit, cp.extraBits

    def decode_file(self, pac_in, wav_out):
        """.pac/.wak -> PCM wav, as the reference Decode pass (pacfile.py:441-499)."""
        huffman = self.new_huffman()
        inFile = self.pacfile.PACFile(pac_in)
        outFile = self.pcmfile.PCMFile(wav_out)
        cp = inFile.OpenForReading()
        cp.bitsPerSample = 16
        cp.nTableIDBits = 4
        outFile.OpenForWriting(cp)
        first = True
        while True:
            data = inFile.ReadDataBlock(cp, huffman)
            if not data:
                break
            if first:
                first = False
                continue
            outFile.WriteDataBlock(data, cp)
        inFile.Close(cp)
        outFile.Close(cp)
