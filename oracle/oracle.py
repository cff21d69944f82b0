"""
oracle.py -- ctypes front end of the CPU oracle (oracle/pac_oracle.c).

TEST INFRASTRUCTURE ONLY.  Allowed importers: tests/, __graft_entry__.smoke(), and
bench.py's cpu_baseline / --impl reference legs.  The product path must never import this.
"""
import ctypes as C
import os
import pickle
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
BUILD_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(BUILD_DIR, "libpacoracle.so")
SRC = os.path.join(HERE, "pac_oracle.c")
HDR = os.path.join(HERE, "pac_oracle.h")
PICKLE = os.path.join(REPO, "perceptual-audio-codec_b200", "huffmanTables.pickle")

NTABLES = 10


def build(force=False):
    """gcc -O2 -fopenmp -shared oracle/pac_oracle.c -> oracle/_build/libpacoracle.so"""
    os.makedirs(BUILD_DIR, exist_ok=True)
    if (not force and os.path.exists(LIB_PATH)
            and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(SRC), os.path.getmtime(HDR))):
        return LIB_PATH
    cmd = ["gcc", "-O2", "-std=gnu99", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-o", LIB_PATH, SRC, "-lm"]
    subprocess.check_call(cmd)
    return LIB_PATH


class OrcParams(C.Structure):
    _fields_ = [("sampleRate", C.c_int32), ("nChannels", C.c_int32), ("nMDCTLines", C.c_int32),
                ("nScaleBits", C.c_int32), ("nMantSizeBits", C.c_int32), ("nTableIDBits", C.c_int32),
                ("targetBitsPerSample", C.c_double), ("window", C.c_int32), ("reserved", C.c_int32)]


class OrcHuff(C.Structure):
    _fields_ = [("nkeys", C.c_int32 * NTABLES), ("off", C.c_int32 * NTABLES),
                ("code", C.POINTER(C.c_uint32)), ("len", C.POINTER(C.c_uint8)),
                ("esc_code", C.c_uint32 * NTABLES), ("esc_len", C.c_int32 * NTABLES)]


class OrcTrace(C.Structure):
    _fields_ = [("lrms", C.POINTER(C.c_int32)), ("oscale", C.POINTER(C.c_int32)),
                ("ba", C.POINTER(C.c_int32)), ("sf", C.POINTER(C.c_int32)),
                ("tableID", C.POINTER(C.c_int32)), ("nbytes", C.POINTER(C.c_int32)),
                ("extraBits", C.POINTER(C.c_int64)), ("bitDeposit", C.POINTER(C.c_int64)),
                ("smr", C.POINTER(C.c_double)), ("lines", C.POINTER(C.c_double)),
                ("mant", C.POINTER(C.c_int32))]


class _StubTable:   # stands in for Huffman.HuffmanTable when unpickling (Huffman.py:138-153)
    pass


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "Huffman":
            return _StubTable
        return super().find_class(module, name)


def load_tables(path=PICKLE):
    """{tableID: {magnitude -> '0101'}} straight from the unchanged pickle."""
    with open(path, "rb") as f:
        tabs = _Unpickler(f, encoding="latin1").load()
    return {int(k): dict(v.encodingTable) for k, v in tabs.items()}


def flatten_tables(tables):
    nkeys, off, codes, lens, esc_code, esc_len = [], [], [], [], [], []
    o = 0
    for tid in range(1, NTABLES + 1):
        enc = tables[tid]
        mk = max(k for k in enc if k >= 0)
        c = np.zeros(mk + 1, dtype=np.uint32)
        l = np.zeros(mk + 1, dtype=np.uint8)
        for k, s in enc.items():
            if k < 0:
                continue
            c[k] = int(s, 2)
            l[k] = len(s)
        nkeys.append(mk + 1)
        off.append(o)
        o += mk + 1
        codes.append(c)
        lens.append(l)
        esc_code.append(int(enc[-1], 2))
        esc_len.append(len(enc[-1]))
    return (np.array(nkeys, np.int32), np.array(off, np.int32), np.concatenate(codes),
            np.concatenate(lens), np.array(esc_code, np.uint32), np.array(esc_len, np.int32))


def default_params(sampleRate=44100, target=2.27, window=0):
    return OrcParams(sampleRate, 2, 1024, 4, 4, 4, target, window, 0)


def read_wav(path):
    """(sampleRate, pcm int16 [n][2]) with the header walk of pcmfile.py:32-57 (4-byte-step scan)."""
    raw = open(path, "rb").read()
    if raw[0:4] != b"RIFF" or raw[8:12] != b"WAVE":
        raise Exception("not a RIFF/WAVE file")
    p = 12
    while raw[p:p + 4] != b"fmt ":
        p += 4
        if p + 4 > len(raw):
            raise Exception("no fmt chunk")
    p += 4
    fsize, tag, nch, rate, bps, align, bits = struct.unpack("<LHHLLHH", raw[p:p + 20])
    p += 20
    if tag != 1 or bits != 16 or nch != 2:
        raise Exception("only 16-bit stereo PCM")
    while raw[p:p + 4] != b"data":
        p += 4
        if p + 4 > len(raw):
            raise Exception("no data chunk")
    nbytes = struct.unpack("<L", raw[p + 4:p + 8])[0]
    n = nbytes // 4
    body = raw[p + 8:p + 8 + n * 4]
    got = len(body) // 4
    pcm = np.zeros((n, 2), dtype=np.int16)
    pcm[:got] = np.frombuffer(body[:got * 4], dtype="<i2").reshape(-1, 2)
    return rate, pcm


def wav_bytes(pcm, sampleRate, numSamplesHdr):
    """pcmfile.py:103-147: header built from cp.numSamples, payload = what was decoded."""
    dataBytes = int(numSamplesHdr) * 2 * 2
    hdr = struct.pack("<4sL4s4sLHHLLHH4sL", b"RIFF", 36 + dataBytes, b"WAVE", b"fmt ", 16, 1, 2,
                      sampleRate, sampleRate * 4, 4, 16, b"data", dataBytes)
    return hdr + np.ascontiguousarray(pcm, dtype="<i2").tobytes()


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class Oracle:
    def __init__(self, pickle_path=PICKLE):
        self.lib = C.CDLL(build())
        L = self.lib
        self.tables = load_tables(pickle_path)
        (self._nkeys, self._off, self._code, self._len, self._esc_code, self._esc_len) = flatten_tables(self.tables)
        h = OrcHuff()
        for i in range(NTABLES):
            h.nkeys[i] = int(self._nkeys[i]); h.off[i] = int(self._off[i])
            h.esc_code[i] = int(self._esc_code[i]); h.esc_len[i] = int(self._esc_len[i])
        h.code = _ptr(self._code, C.c_uint32)
        h.len = _ptr(self._len, C.c_uint8)
        self.huff = h
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        L.orc_band_layout.argtypes = [C.c_int, C.c_int, ip]
        L.orc_sine_window.argtypes = [dp, C.c_int]
        L.orc_hann_window.argtypes = [dp, C.c_int]
        L.orc_kbd_window.argtypes = [dp, dp, C.c_int, C.c_double]
        L.orc_fft.argtypes = [dp, dp, C.c_int, C.c_int]
        L.orc_mdct.argtypes = [dp, C.c_int, C.c_int, dp]
        L.orc_imdct.argtypes = [dp, C.c_int, C.c_int, dp]
        for f in (L.orc_spl, L.orc_intensity, L.orc_thresh, L.orc_bark):
            f.argtypes = [C.c_double]; f.restype = C.c_double
        L.orc_quantize_uniform.argtypes = [C.c_double, C.c_int]
        L.orc_vquantize_uniform.argtypes = [dp, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.orc_vdequantize_uniform.argtypes = [C.POINTER(C.c_uint64), C.c_int, C.c_int, dp]
        L.orc_scale_factor.argtypes = [C.c_double, C.c_int, C.c_int]
        L.orc_vmantissa.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.orc_vdequantize.argtypes = [C.c_int, C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int, dp]
        L.orc_bitalloc.argtypes = [C.c_double, C.c_int64, C.c_int, C.c_int, ip, dp, ip, ip, C.POINTER(C.c_int64)]
        L.orc_bitalloc_alt.argtypes = [C.c_int, C.c_double, C.c_int, C.c_int, ip, dp, ip]
        L.orc_bitalloc_alt.restype = C.c_int
        L.orc_calc_bthr.argtypes = [dp, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.orc_calc_smrs.argtypes = [dp, C.c_int, dp, C.c_int, C.c_int, C.c_int, ip, C.c_int, dp, dp]
        L.orc_stereo_smr.argtypes = [dp, dp, C.c_int, dp, dp, C.c_int, ip, C.c_int, ip, C.c_int, ip, dp, dp, dp]
        L.orc_lrms.argtypes = [dp, dp, C.c_int, ip, C.c_int, ip]
        L.orc_encode_stream.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcHuff), C.POINTER(C.c_int16), C.c_int64,
                                        C.POINTER(C.c_uint8), C.c_int64, C.POINTER(OrcTrace), C.POINTER(C.c_int64)]
        L.orc_encode_stream.restype = C.c_int64
        L.orc_encoded_blocks.argtypes = [C.c_int64, C.c_int]
        L.orc_encoded_blocks.restype = C.c_int64
        L.orc_decode_stream.argtypes = [C.POINTER(OrcHuff), C.POINTER(C.c_uint8), C.c_int64, C.POINTER(C.c_int16),
                                        C.c_int64, C.POINTER(OrcParams), C.POINTER(C.c_int64)]
        L.orc_decode_stream.restype = C.c_int64
        L.orc_decode_stream_w.argtypes = [C.POINTER(OrcHuff), C.POINTER(C.c_uint8), C.c_int64, C.POINTER(C.c_int16),
                                          C.c_int64, C.POINTER(OrcParams), C.POINTER(C.c_int64), C.c_int]
        L.orc_decode_stream_w.restype = C.c_int64
        L.orc_encode_batch.argtypes = [C.POINTER(OrcParams), C.POINTER(OrcHuff), C.POINTER(C.c_int16), C.c_int64, C.c_int,
                                       C.POINTER(C.c_uint8), C.c_int64, C.POINTER(C.c_int64), C.c_int]

    # ---- small wrappers -------------------------------------------------
    def band_layout(self, nMDCTLines=1024, sampleRate=44100):
        n = np.zeros(25, dtype=np.int32)
        self.lib.orc_band_layout(nMDCTLines, sampleRate, _ptr(n, C.c_int32))
        return n

    def sine_window(self, x):
        x = np.array(x, dtype=np.float64); self.lib.orc_sine_window(_ptr(x, C.c_double), len(x)); return x

    def hann_window(self, x):
        x = np.array(x, dtype=np.float64); self.lib.orc_hann_window(_ptr(x, C.c_double), len(x)); return x

    def kbd_window(self, x, alpha=4.0):
        x = np.array(x, dtype=np.float64); o = np.empty_like(x)
        self.lib.orc_kbd_window(_ptr(x, C.c_double), _ptr(o, C.c_double), len(x), alpha); return o

    def mdct(self, x, a, b):
        x = np.array(x, dtype=np.float64); X = np.empty((a + b) // 2)
        self.lib.orc_mdct(_ptr(x, C.c_double), a, b, _ptr(X, C.c_double)); return X

    def imdct(self, X, a, b):
        X = np.array(X, dtype=np.float64); x = np.empty(a + b)
        self.lib.orc_imdct(_ptr(X, C.c_double), a, b, _ptr(x, C.c_double)); return x

    def vquantize_uniform(self, x, nBits):
        x = np.array(x, dtype=np.float64); q = np.empty(len(x), dtype=np.uint64)
        self.lib.orc_vquantize_uniform(_ptr(x, C.c_double), len(x), nBits, _ptr(q, C.c_uint64)); return q

    def vdequantize_uniform(self, q, nBits):
        q = np.array(q, dtype=np.uint64); x = np.empty(len(q))
        self.lib.orc_vdequantize_uniform(_ptr(q, C.c_uint64), len(q), nBits, _ptr(x, C.c_double)); return x

    def scale_factor(self, x, nScaleBits=3, nMantBits=5):
        return int(self.lib.orc_scale_factor(float(x), nScaleBits, nMantBits))

    def vmantissa(self, x, scale, nScaleBits=3, nMantBits=5):
        x = np.array(x, dtype=np.float64); m = np.empty(len(x), dtype=np.uint64)
        self.lib.orc_vmantissa(_ptr(x, C.c_double), len(x), scale, nScaleBits, nMantBits, _ptr(m, C.c_uint64)); return m

    def vdequantize(self, scale, m, nScaleBits=3, nMantBits=5):
        m = np.array(m, dtype=np.int64); x = np.empty(len(m))
        self.lib.orc_vdequantize(scale, _ptr(m, C.c_int64), len(m), nScaleBits, nMantBits, _ptr(x, C.c_double)); return x

    def bitalloc(self, bitBudget, extraBits, maxMantBits, nBands, nLines, SMR, LRMS):
        nLines = np.array(nLines, dtype=np.int32); SMR = np.array(SMR, dtype=np.float64)
        LRMS = np.array(LRMS, dtype=np.int32); bits = np.zeros(nBands, dtype=np.int32); d = C.c_int64()
        self.lib.orc_bitalloc(bitBudget, int(extraBits), maxMantBits, nBands, _ptr(nLines, C.c_int32),
                              _ptr(SMR, C.c_double), _ptr(LRMS, C.c_int32), _ptr(bits, C.c_int32), C.byref(d))
        return bits, int(d.value)

    def bitalloc_alt(self, mode, bitBudget, maxMantBits, nBands, nLines, level=None):
        """bitalloc.py:22-125; mode 'uniform' | 'constsnr' | 'constmnr'.  Raises where the reference would not terminate."""
        m = {"uniform": 0, "constsnr": 1, "constmnr": 2}[mode]
        nLines = np.array(nLines, dtype=np.int32); bits = np.zeros(nBands, dtype=np.int32)
        lv = np.zeros(nBands) if level is None else np.ascontiguousarray(np.broadcast_to(np.asarray(level, dtype=np.float64), (nBands,)))
        if self.lib.orc_bitalloc_alt(m, float(bitBudget), maxMantBits, nBands, _ptr(nLines, C.c_int32), _ptr(lv, C.c_double), _ptr(bits, C.c_int32)):
            raise RuntimeError("the reference allocator does not terminate for this input")
        return bits

    def calc_smrs(self, data, mdct, scale, sampleRate, nLines):
        data = np.array(data, dtype=np.float64); mdct = np.array(mdct, dtype=np.float64)
        nLines = np.array(nLines, dtype=np.int32); smr = np.empty(len(nLines)); thr = np.empty(len(mdct))
        self.lib.orc_calc_smrs(_ptr(data, C.c_double), len(data), _ptr(mdct, C.c_double), len(mdct), scale, sampleRate,
                               _ptr(nLines, C.c_int32), len(nLines), _ptr(smr, C.c_double), _ptr(thr, C.c_double))
        return smr, thr

    def lrms(self, l, r, nLines):
        l = np.array(l, dtype=np.float64); r = np.array(r, dtype=np.float64)
        nLines = np.array(nLines, dtype=np.int32); out = np.zeros(len(nLines), dtype=np.int32)
        self.lib.orc_lrms(_ptr(l, C.c_double), _ptr(r, C.c_double), len(l), _ptr(nLines, C.c_int32), len(nLines), _ptr(out, C.c_int32))
        return out

    def stereo_smr(self, d0, d1, X0, X1, scale, sampleRate, nLines, LRMS):
        """d0,d1: SINE-WINDOWED time blocks (psychoac.py:506 receives them that way, codec.py:239-240)."""
        d0 = np.array(d0, dtype=np.float64); d1 = np.array(d1, dtype=np.float64)
        X0 = np.array(X0, dtype=np.float64); X1 = np.array(X1, dtype=np.float64)
        nLines = np.array(nLines, dtype=np.int32); LRMS = np.array(LRMS, dtype=np.int32)
        scale = np.array(scale, dtype=np.int32); nB = len(nLines); nL = len(X0)
        smr = np.empty((2, nB)); lines = np.empty((2, nL)); bthr = np.empty((6, nL))
        self.lib.orc_stereo_smr(_ptr(d0, C.c_double), _ptr(d1, C.c_double), len(d0), _ptr(X0, C.c_double), _ptr(X1, C.c_double),
                                nL, _ptr(scale, C.c_int32), sampleRate, _ptr(nLines, C.c_int32), nB, _ptr(LRMS, C.c_int32),
                                _ptr(smr, C.c_double), _ptr(lines, C.c_double), _ptr(bthr, C.c_double))
        return smr, lines, bthr

    # ---- streams ----------------------------------------------------------
    def encoded_blocks(self, nSamples, nMDCTLines=1024):
        return int(self.lib.orc_encoded_blocks(int(nSamples), nMDCTLines))

    def encode_stream(self, pcm, params=None, trace=False, cap=None):
        """pcm int16 [n][2] -> (bytes, trace dict | None, (bitDeposit, extraBits))"""
        p = params or default_params()
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        n = pcm.shape[0]
        nb = self.encoded_blocks(n, p.nMDCTLines)
        if cap is None:
            cap = 128 + nb * 2 * (4 + 4096)
        out = np.zeros(cap, dtype=np.uint8)
        fs = np.zeros(2, dtype=np.int64)
        tr = None
        trs = None
        if trace:
            nB = 25; hN = p.nMDCTLines
            tr = {"lrms": np.zeros(nb, np.int32), "oscale": np.zeros((nb, 2), np.int32),
                  "ba": np.zeros((nb, 2, nB), np.int32), "sf": np.zeros((nb, 2, nB), np.int32),
                  "tableID": np.zeros((nb, 2), np.int32), "nbytes": np.zeros((nb, 2), np.int32),
                  "extraBits": np.zeros(nb, np.int64), "bitDeposit": np.zeros(nb, np.int64),
                  "smr": np.zeros((nb, 2, nB)), "lines": np.zeros((nb, 2, hN)),
                  "mant": np.zeros((nb, 2, hN), np.int32)}
            trs = OrcTrace(_ptr(tr["lrms"], C.c_int32), _ptr(tr["oscale"], C.c_int32), _ptr(tr["ba"], C.c_int32),
                           _ptr(tr["sf"], C.c_int32), _ptr(tr["tableID"], C.c_int32), _ptr(tr["nbytes"], C.c_int32),
                           _ptr(tr["extraBits"], C.c_int64), _ptr(tr["bitDeposit"], C.c_int64),
                           _ptr(tr["smr"], C.c_double), _ptr(tr["lines"], C.c_double), _ptr(tr["mant"], C.c_int32))
        r = self.lib.orc_encode_stream(C.byref(p), C.byref(self.huff), _ptr(pcm, C.c_int16), n, _ptr(out, C.c_uint8), cap,
                                       C.byref(trs) if trs is not None else None, _ptr(fs, C.c_int64))
        if r < 0:
            raise RuntimeError("orc_encode_stream failed: %d" % r)
        return out[:r].tobytes(), tr, (int(fs[0]), int(fs[1]))

    def decode_stream(self, pac, window=0):
        """pac bytes -> (pcm int16 [n][2], sampleRate, numSamples from the header)"""
        buf = np.frombuffer(pac, dtype=np.uint8).copy()
        nblocks_upper = len(buf) // 8 + 2
        cap = nblocks_upper * 1024 if len(buf) < (1 << 20) else (len(buf) // 200 + 2) * 1024
        pcm = np.zeros((cap, 2), dtype=np.int16)
        hdr = OrcParams()
        ns = C.c_int64()
        r = self.lib.orc_decode_stream_w(C.byref(self.huff), _ptr(buf, C.c_uint8), len(buf), _ptr(pcm, C.c_int16), cap,
                                         C.byref(hdr), C.byref(ns), int(window))
        if r < 0:
            raise RuntimeError("orc_decode_stream failed: %d" % r)
        return pcm[:r].copy(), int(hdr.sampleRate), int(ns.value)

    def encode_batch(self, pcm, params=None, nthreads=1, cap=None):
        """pcm int16 [S][n][2] -> list of bytes; runs on `nthreads` host threads (CPU baseline)."""
        p = params or default_params()
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        S, n = pcm.shape[0], pcm.shape[1]
        nb = self.encoded_blocks(n, p.nMDCTLines)
        if cap is None:
            cap = 128 + nb * 2 * (4 + 2048)
        out = np.zeros((S, cap), dtype=np.uint8)
        nbytes = np.zeros(S, dtype=np.int64)
        rc = self.lib.orc_encode_batch(C.byref(p), C.byref(self.huff), _ptr(pcm, C.c_int16), n, S, _ptr(out, C.c_uint8),
                                       cap, _ptr(nbytes, C.c_int64), nthreads)
        if rc < 0:
            raise RuntimeError("orc_encode_batch failed: %d" % rc)
        return [out[s, :nbytes[s]].tobytes() for s in range(S)]

    def encode_wav(self, path, params=None, trace=False):
        rate, pcm = read_wav(path)
        p = params or default_params(rate)
        return self.encode_stream(pcm, p, trace)

    def decode_to_wav_bytes(self, pac):
        pcm, rate, ns = self.decode_stream(pac)
        return wav_bytes(pcm, rate, ns)


_singleton = None


def get():
    global _singleton
    if _singleton is None:
        _singleton = Oracle()
    return _singleton


# ---------------------------------------------------------------- Huffman trainer (Huffman.py:27-250), test infrastructure
def train_huffman(chunks, low_freq=10, escape_code=-1, state=None):
    """CPU restatement of HuffmanTrainer.countFreq (once per chunk) + constructHuffmanTable: returns (encodingTable, state).
    `state` carries what the reference keeps in CLASS attributes from one trainer to the next in a process (Huffman.py:30-34,
    158-160): the statistics dict, the class-level deque -- makeHuffmanNodeQueue appends the new leaves to it and then binds
    a sorted COPY to the instance (:103-108), so the leaves of earlier trainers are still in it and join the next tree --
    and the code table dict, which is never cleared.  Otherwise literal: dict in first-insertion order (:74-78), stable
    sort by frequency (:94), escape weight = number of rare codes (:101), queue re-sorted after every join (:118-119),
    zero = first popped (:226), depth-first zero-then-one table walk (:238-241, a repeated code keeps the last leaf visited)."""
    if state is None:
        state = {"statistics": {}, "queue": [], "table": {}}
    stats = state["statistics"]
    for chunk in chunks:
        for c in chunk:
            c = int(c)
            stats[c] = stats.get(c, 0) + 1
    esc = 0
    for code, freq in sorted(stats.items(), key=lambda t: t[1]):
        if freq < low_freq:
            esc += 1
        else:
            state["queue"].append([freq, code, None, None])
    state["queue"].append([esc, escape_code, None, None])
    queue = sorted(state["queue"], key=lambda t: t[0])
    while len(queue) > 1:
        a, b = queue.pop(0), queue.pop(0)
        queue.append([a[0] + b[0], None, a, b])
        queue.sort(key=lambda t: t[0])
    table = state["table"]
    stack = [(queue[0], "")]
    while stack:
        node, code = stack.pop()
        if node[1] is not None:
            table[node[1]] = code
            continue
        stack.append((node[3], code + "1"))
        stack.append((node[2], code + "0"))
    return dict(table), state
