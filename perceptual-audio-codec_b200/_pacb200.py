"""
_pacb200.py -- ctypes binding of libpacb200.so (include/pac_b200.h) and the Engine object the reference-named
shim modules (pacfile.py, codec.py, psychoac.py, mdct.py, window.py, quantize.py, bitalloc.py) share.

There is no CPU fallback: if the CUDA library is missing or no sm_100 device is visible, importing the engine
raises.  Build the library with `python __graft_entry__.py` (or perceptual-audio-codec_b200/build.py).
"""
import ctypes as C
import os
import pickle

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpacb200.so")

PAC_PRECISION_FP64 = 0
PAC_PRECISION_FP32 = 1
NTABLES = 10
MAX_BANDS = 32

ERRORS = {-1: "PAC_E_ARG", -2: "PAC_E_CUDA", -3: "PAC_E_OVERFLOW", -4: "PAC_E_FORMAT", -5: "PAC_E_NODEVICE"}


class PacError(Exception):
    def __init__(self, code, msg):
        Exception.__init__(self, "%s (%d): %s" % (ERRORS.get(code, "PAC_E_?"), code, msg))
        self.code = code


class PacParams(C.Structure):
    _fields_ = [("sampleRate", C.c_int32), ("nChannels", C.c_int32), ("nMDCTLines", C.c_int32),
                ("nScaleBits", C.c_int32), ("nMantSizeBits", C.c_int32), ("nTableIDBits", C.c_int32),
                ("targetBitsPerSample", C.c_double), ("window", C.c_int32), ("reserved", C.c_int32)]


class PacHuffTables(C.Structure):
    _fields_ = [("nkeys", C.c_int32 * NTABLES), ("off", C.c_int32 * NTABLES),
                ("code", C.POINTER(C.c_uint32)), ("len", C.POINTER(C.c_uint8)),
                ("esc_code", C.c_uint32 * NTABLES), ("esc_len", C.c_int32 * NTABLES)]


class PacTrace(C.Structure):
    _fields_ = [("lrms", C.POINTER(C.c_int32)), ("oscale", C.POINTER(C.c_int32)), ("smr", C.POINTER(C.c_double)),
                ("lines", C.POINTER(C.c_double)), ("ba", C.POINTER(C.c_int32)), ("sf", C.POINTER(C.c_int32)),
                ("tableID", C.POINTER(C.c_int32)), ("nbytes", C.POINTER(C.c_int32)),
                ("extraBits", C.POINTER(C.c_int64)), ("bitDeposit", C.POINTER(C.c_int64)), ("mant", C.POINTER(C.c_int32))]


class PacStreamState(C.Structure):
    _fields_ = [("extraBits", C.c_int64), ("bitDeposit", C.c_int64)]


_lib = None


def lib():
    """Load libpacb200.so; fail loudly when it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libpacb200.so not found at %s: build it with `python __graft_entry__.py` "
                          "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32p, i64p, dp, u8p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
    L.pac_ctx_create.argtypes = [C.c_int, C.c_int, C.POINTER(PacParams), C.POINTER(PacHuffTables), C.POINTER(vp)]
    L.pac_ctx_destroy.argtypes = [vp]
    L.pac_ctx_destroy.restype = None
    L.pac_last_error.argtypes = [vp]
    L.pac_last_error.restype = C.c_char_p
    L.pac_version.restype = C.c_char_p
    L.pac_pinned_alloc.argtypes = [C.c_size_t]
    L.pac_pinned_alloc.restype = vp
    L.pac_pinned_free.argtypes = [vp]
    L.pac_pinned_free.restype = None
    L.pac_h2d_bandwidth.argtypes = [vp, C.c_size_t, C.c_int, dp]
    L.pac_band_layout.argtypes = [vp, i32p, i32p]
    L.pac_launch_count.argtypes = [vp]
    L.pac_launch_count.restype = C.c_int64
    L.pac_set_stream.argtypes = [vp, vp]
    L.pac_timing_enable.argtypes = [vp, C.c_int]
    L.pac_timing_get.argtypes = [vp, dp, i64p]
    L.pac_num_blocks.argtypes = [vp, C.c_int64]
    L.pac_num_blocks.restype = C.c_int64
    L.pac_encode_bound.argtypes = [vp, C.c_int64]
    L.pac_encode_bound.restype = C.c_int64
    L.pac_decode_bound.argtypes = [vp, C.c_int64]
    L.pac_decode_bound.restype = C.c_int64
    L.pac_encode_batch.argtypes = [vp, vp, C.c_int64, i64p, C.c_int, vp, C.c_int64, i64p, i64p, C.POINTER(PacTrace)]
    L.pac_mdct_batch.argtypes = [vp, vp, C.c_int64, i64p, C.c_int, dp, i32p, dp]
    L.pac_analysis_batch.argtypes = [vp, vp, C.c_int64, i64p, C.c_int, dp]
    L.pac_decode_batch.argtypes = [vp, vp, i64p, C.c_int, vp, C.c_int64, i64p, i64p, i32p]
    L.pac_decode_batch_strided.argtypes = [vp, vp, i64p, i64p, C.c_int, vp, C.c_int64, i64p, i64p, i32p]
    L.pac_encode_blocks.argtypes = [vp, dp, C.c_int, C.POINTER(PacStreamState), i32p, i32p, i32p, i32p, i32p, i32p, u8p, C.c_int64, i32p]
    L.pac_decode_blocks.argtypes = [vp, i32p, i32p, i32p, i32p, i32p, C.c_int, dp]
    L.pac_unpack_blocks.argtypes = [vp, u8p, C.c_int64, i32p, C.c_int, i32p, i32p, i32p, i32p, i32p, i32p]
    L.pac_window.argtypes = [vp, C.c_int, dp, C.c_int, C.c_int]
    L.pac_mdct.argtypes = [vp, dp, C.c_int, C.c_int, dp]
    L.pac_imdct.argtypes = [vp, dp, C.c_int, C.c_int, dp]
    L.pac_analysis.argtypes = [vp, dp, C.c_int, i32p, i32p, dp, dp, dp, dp]
    L.pac_calc_smrs.argtypes = [vp, dp, dp, C.c_int, C.c_int, dp]
    L.pac_masked_threshold.argtypes = [vp, dp, C.c_int, C.c_int, dp]
    L.pac_huffman_select.argtypes = [vp, C.POINTER(C.c_uint32), i32p, C.c_int, i32p, i64p]
    L.pac_bitalloc.argtypes = [vp, C.c_int, dp, i64p, C.c_int, dp, i32p, i32p, i64p]
    L.pac_bitalloc_alt.argtypes = [vp, C.c_int, C.c_int, dp, C.c_int, dp, i32p]
    L.pac_histogram.argtypes = [vp, C.c_void_p, C.c_int64, C.c_int64, C.c_int, i64p, i64p]
    L.pac_scale_factor.argtypes = [vp, dp, C.c_int, C.c_int, C.c_int, i32p]
    L.pac_vquantize_uniform.argtypes = [vp, dp, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    L.pac_vdequantize_uniform.argtypes = [vp, C.POINTER(C.c_uint64), C.c_int, C.c_int, dp]
    L.pac_vmantissa.argtypes = [vp, dp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    L.pac_vdequantize.argtypes = [vp, C.c_int, i64p, C.c_int, C.c_int, C.c_int, dp]
    _lib = L
    return L


# ------------------------------------------------------------------ pickled tables (unchanged reference fixture)

class _StubTable(object):
    pass


class _Unpickler(pickle.Unpickler):
    """Restricted unpickler for the two reference fixtures.  They reference exactly three globals (Huffman.HuffmanTable,
    Huffman.Histogram, collections.deque); anything else in a pickle found in the working directory is refused instead of
    imported -- a hostile huffmanTables.pickle must not be able to run code in a process that merely creates an Engine."""
    classes = {}

    def find_class(self, module, name):
        if module == "Huffman" and name in ("HuffmanTable", "Histogram", "HuffmanNode"):
            return self.classes.get(name, _StubTable)
        if (module, name) == ("collections", "deque"):
            import collections
            return collections.deque
        # protocol-0 pickles of new-style classes (what the trainer shim writes back under Python 3, Huffman.py:195-211) are rebuilt
        # through copy_reg._reconstructor(cls, object, None); cls itself still has to pass the whitelist above
        if (module, name) in (("copy_reg", "_reconstructor"), ("copyreg", "_reconstructor")):
            import copyreg
            return copyreg._reconstructor
        if (module, name) in (("__builtin__", "object"), ("builtins", "object")):
            return object
        raise pickle.UnpicklingError("refusing to unpickle %s.%s: the codec's table fixtures only hold Huffman.HuffmanTable, "
                                     "Huffman.Histogram, Huffman.HuffmanNode and collections.deque" % (module, name))


def safe_load(handle, classes=None):
    """pickle.load for huffmanTables.pickle / histograms.pickle with the global whitelist above; `classes` maps the
    fixture's class names to the caller's classes (the Huffman shim passes its own HuffmanTable / Histogram)."""
    u = _Unpickler(handle, encoding="latin1")
    u.classes = classes or {}
    return u.load()


def find_pickle(name="huffmanTables.pickle"):
    """The reference opens its pickles relative to the CWD (Huffman.py:257-260); fall back to this directory."""
    for d in (os.getcwd(), HERE):
        p = os.path.join(d, name)
        if os.path.exists(p):
            return p
    raise IOError("cannot find %s in the working directory or in %s" % (name, HERE))


def load_encoding_tables(path=None):
    with open(path or find_pickle(), "rb") as f:
        tabs = _Unpickler(f, encoding="latin1").load()
    return {int(k): dict(v.encodingTable) for k, v in tabs.items()}


def flatten_tables(tables):
    """{id: {magnitude: '0101', -1: escape}} -> arrays of PacHuffTables."""
    nkeys, off, codes, lens, esc_code, esc_len = [], [], [], [], [], []
    o = 0
    for tid in range(1, NTABLES + 1):
        enc = tables[tid]
        mk = max(k for k in enc if k >= 0)
        c = np.zeros(mk + 1, dtype=np.uint32)
        l = np.zeros(mk + 1, dtype=np.uint8)
        for k, s in enc.items():
            if k >= 0:
                c[k] = int(s, 2)
                l[k] = len(s)
        nkeys.append(mk + 1); off.append(o); o += mk + 1
        codes.append(c); lens.append(l)
        esc_code.append(int(enc[-1], 2)); esc_len.append(len(enc[-1]))
    return (np.array(nkeys, np.int32), np.array(off, np.int32), np.ascontiguousarray(np.concatenate(codes)),
            np.ascontiguousarray(np.concatenate(lens)), np.array(esc_code, np.uint32), np.array(esc_len, np.int32))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _vp(a):
    """raw address of a numpy array or of anything exposing data_ptr() (e.g. a CUDA torch tensor)"""
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


class _PinnedBlock(object):
    """owner of one pac_pinned_alloc allocation (freed when the last numpy view dies)"""

    def __init__(self, nbytes):
        self.nbytes = int(nbytes)
        self.ptr = lib().pac_pinned_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError("pac_pinned_alloc(%d) failed" % self.nbytes)
        self.__array_interface__ = {"shape": (max(self.nbytes, 1),), "typestr": "|u1", "data": (self.ptr, False), "version": 3}

    def __del__(self):
        try:
            if self.ptr:
                lib().pac_pinned_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype=np.uint8):
    """numpy array in page-locked, device-mapped host memory (pac_pinned_alloc): the buffers of encode_batch / decode_batch then move
    by DMA, and a pinned `out` of encode_batch is written in place by the pack kernel."""
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    blk = _PinnedBlock(n)
    return np.asarray(blk)[:n].view(dt).reshape(shape)          # the view keeps blk alive through its base chain


def h2d_bandwidth(arr, device=0):
    """measured host -> device copy bandwidth (GB/s) of a numpy array's memory (pac_h2d_bandwidth)"""
    g = C.c_double(0.0)
    rc = lib().pac_h2d_bandwidth(C.c_void_p(arr.ctypes.data), arr.nbytes, int(device), C.byref(g))
    if rc:
        raise PacError(rc, "pac_h2d_bandwidth failed")
    return float(g.value)


class Engine(object):
    """One context on one GPU (PacCtx).  precision: 'fp64' (verification, bit-exact) or 'fp32' (fast)."""

    def __init__(self, device=0, precision="fp64", sampleRate=44100, nMDCTLines=1024, nScaleBits=4, nMantSizeBits=4,
                 nTableIDBits=4, targetBitsPerSample=2.27, tables=None, window="sine"):
        L = lib()
        self.tables = tables if tables is not None else load_encoding_tables()
        (self._nkeys, self._off, self._code, self._len, self._esc_code, self._esc_len) = flatten_tables(self.tables)
        h = PacHuffTables()
        for i in range(NTABLES):
            h.nkeys[i] = int(self._nkeys[i]); h.off[i] = int(self._off[i])
            h.esc_code[i] = int(self._esc_code[i]); h.esc_len[i] = int(self._esc_len[i])
        h.code = _p(self._code, C.c_uint32)
        h.len = _p(self._len, C.c_uint8)
        self.params = PacParams(int(sampleRate), 2, int(nMDCTLines), int(nScaleBits), int(nMantSizeBits), int(nTableIDBits),
                                float(targetBitsPerSample), {"sine": 0, "kbd": 1}[window], 0)
        self.precision = {"fp64": PAC_PRECISION_FP64, "fp32": PAC_PRECISION_FP32}[precision]
        self.precision_name = precision
        self.device = device
        ctx = C.c_void_p()
        rc = L.pac_ctx_create(device, self.precision, C.byref(self.params), C.byref(h), C.byref(ctx))
        if rc:
            raise PacError(rc, L.pac_last_error(None).decode())
        self.ctx = ctx
        nl = np.zeros(MAX_BANDS, np.int32)
        nb = C.c_int32()
        self._ck(L.pac_band_layout(ctx, _p(nl, C.c_int32), C.byref(nb)))
        self.nBands = nb.value
        self.nLines = nl[:nb.value].copy()
        self.M = int(nMDCTLines)
        self.N = 2 * self.M

    def close(self):
        if getattr(self, "ctx", None):
            lib().pac_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise PacError(rc, lib().pac_last_error(self.ctx).decode())

    @property
    def launches(self):
        return int(lib().pac_launch_count(self.ctx))

    KINDS = ("analysis", "scan", "pack", "index", "unpack", "synth", "mdct")

    def set_stream(self, cuda_stream_handle):
        """use the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream).  Handle 0 is CUDA's legacy default
        stream (torch's default stream reports 0) and is used as such; None restores the context's own stream
        (PAC_STREAM_OWN)."""
        h = C.c_void_p(-1) if cuda_stream_handle is None else C.c_void_p(int(cuda_stream_handle))
        self._ck(lib().pac_set_stream(self.ctx, h))

    def timing(self, on=True):
        self._ck(lib().pac_timing_enable(self.ctx, 1 if on else 0))

    def timing_get(self):
        """{kernel kind: (total device ms, launches)} since timing(True)"""
        ms = np.zeros(8); cnt = np.zeros(8, np.int64)
        self._ck(lib().pac_timing_get(self.ctx, _p(ms, C.c_double), _p(cnt, C.c_int64)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.KINDS)}

    # ---------------------------------------------------------------- whole streams
    def num_blocks(self, nSamples):
        return int(lib().pac_num_blocks(self.ctx, int(nSamples)))

    def encode_bound(self, nSamples):
        return int(lib().pac_encode_bound(self.ctx, int(nSamples)))

    def encode_batch(self, pcm, nSamples=None, out=None, cap=None, trace=False, raw=False):
        """pcm: int16 [S][n][2] numpy array (host) or CUDA tensor.  Returns list of bytes per stream (host out) or
        (out, outBytes) when `out` is given / raw=True; with trace=True also the dict of per-block taps."""
        S, stride = int(pcm.shape[0]), int(pcm.shape[1])
        if nSamples is None:
            nSamples = np.full(S, stride, dtype=np.int64)
        nSamples = np.ascontiguousarray(nSamples, dtype=np.int64)
        if cap is None:
            cap = self.encode_bound(int(nSamples.max()) if S else 0)
        own_out = out is None
        if own_out:
            out = np.empty((S, cap), dtype=np.uint8)
        outBytes = np.zeros(S, dtype=np.int64)
        final = np.zeros((S, 2), dtype=np.int64)
        tr = None
        trs = None
        if trace:
            B = max(self.num_blocks(int(n)) for n in nSamples)
            NB, M = self.nBands, self.M
            tr = {"lrms": np.zeros((S, B), np.int32), "oscale": np.zeros((S, B, 2), np.int32),
                  "smr": np.zeros((S, B, 2, NB)), "lines": np.zeros((S, B, 2, M)),
                  "ba": np.zeros((S, B, 2, NB), np.int32), "sf": np.zeros((S, B, 2, NB), np.int32),
                  "tableID": np.zeros((S, B, 2), np.int32), "nbytes": np.zeros((S, B, 2), np.int32),
                  "extraBits": np.zeros((S, B), np.int64), "bitDeposit": np.zeros((S, B), np.int64),
                  "mant": np.zeros((S, B, 2, M), np.int32)}
            trs = PacTrace(_p(tr["lrms"], C.c_int32), _p(tr["oscale"], C.c_int32), _p(tr["smr"], C.c_double),
                           _p(tr["lines"], C.c_double), _p(tr["ba"], C.c_int32), _p(tr["sf"], C.c_int32),
                           _p(tr["tableID"], C.c_int32), _p(tr["nbytes"], C.c_int32), _p(tr["extraBits"], C.c_int64),
                           _p(tr["bitDeposit"], C.c_int64), _p(tr["mant"], C.c_int32))
        if isinstance(pcm, np.ndarray):
            pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        rc = lib().pac_encode_batch(self.ctx, _vp(pcm), stride, _p(nSamples, C.c_int64), S, _vp(out), int(cap),
                                    _p(outBytes, C.c_int64), _p(final, C.c_int64), C.byref(trs) if trs is not None else None)
        self._ck(rc)
        self.last_final_state = final
        if own_out and not raw:
            res = [out[s, :outBytes[s]].tobytes() for s in range(S)]
            return (res, tr) if trace else res
        return (out, outBytes, tr) if trace else (out, outBytes)

    def mdct_batch(self, pcm, nSamples=None, want=True):
        """The window + MDCT stage of encode_batch by itself (pac_mdct_batch): int16 [S][n][2] -> (lines [S][B][2][M] scaled L/R
        lines, oscale [S][B][2], device ms of the kernel launches); want=False keeps the results on the device (stage timing)."""
        S, stride = int(pcm.shape[0]), int(pcm.shape[1])
        if nSamples is None:
            nSamples = np.full(S, stride, dtype=np.int64)
        nSamples = np.ascontiguousarray(nSamples, dtype=np.int64)
        B = max(self.num_blocks(int(n)) for n in nSamples)
        lines = np.zeros((S, B, 2, self.M)) if want else None
        osc = np.zeros((S, B, 2), np.int32) if want else None
        ms = C.c_double(0.0)
        if isinstance(pcm, np.ndarray):
            pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        self._ck(lib().pac_mdct_batch(self.ctx, _vp(pcm), stride, _p(nSamples, C.c_int64), S,
                                      _p(lines, C.c_double) if want else None, _p(osc, C.c_int32) if want else None, C.byref(ms)))
        return lines, osc, float(ms.value)

    def analysis_batch_ms(self, pcm, nSamples=None):
        """device ms of the whole analysis kernel run by itself over the batch (pac_analysis_batch; stage timing)"""
        S, stride = int(pcm.shape[0]), int(pcm.shape[1])
        if nSamples is None:
            nSamples = np.full(S, stride, dtype=np.int64)
        nSamples = np.ascontiguousarray(nSamples, dtype=np.int64)
        ms = C.c_double(0.0)
        self._ck(lib().pac_analysis_batch(self.ctx, _vp(pcm), stride, _p(nSamples, C.c_int64), S, C.byref(ms)))
        return float(ms.value)

    def decode_bound(self, nbytes):
        return int(lib().pac_decode_bound(self.ctx, int(nbytes)))

    def decode_batch(self, pacs, stride=None):
        """pacs: list of bytes (.pac file images).  Returns list of (pcm int16 [n][2], sampleRate, numSamplesHdr)."""
        S = len(pacs)
        off = np.zeros(S + 1, dtype=np.int64)
        for i, p in enumerate(pacs):
            off[i + 1] = off[i] + len(p)
        blob = np.frombuffer(b"".join(pacs), dtype=np.uint8).copy()
        if stride is None:
            stride = max(self.decode_bound(len(p)) for p in pacs)
        pcm = np.zeros((S, stride, 2), dtype=np.int16)
        ns = np.zeros(S, dtype=np.int64)
        hn = np.zeros(S, dtype=np.int64)
        hr = np.zeros(S, dtype=np.int32)
        self._ck(lib().pac_decode_batch(self.ctx, _vp(blob), _p(off, C.c_int64), S, _vp(pcm), int(stride),
                                        _p(ns, C.c_int64), _p(hn, C.c_int64), _p(hr, C.c_int32)))
        return [(pcm[s, :ns[s]].copy(), int(hr[s]), int(hn[s])) for s in range(S)]

    def decode_batch_strided(self, pac, beg, length, pcm, stride):
        """pac / pcm: numpy arrays or CUDA tensors; stream s = pac[beg[s] : beg[s]+length[s]].  Returns (nSamples, hdrNumSamples, hdrRate)."""
        beg = np.ascontiguousarray(beg, np.int64); length = np.ascontiguousarray(length, np.int64)
        S = len(beg)
        ns = np.zeros(S, np.int64); hn = np.zeros(S, np.int64); hr = np.zeros(S, np.int32)
        self._ck(lib().pac_decode_batch_strided(self.ctx, _vp(pac), _p(beg, C.c_int64), _p(length, C.c_int64), S, _vp(pcm), int(stride),
                                                _p(ns, C.c_int64), _p(hn, C.c_int64), _p(hr, C.c_int32)))
        return ns, hn, hr

    # ---------------------------------------------------------------- per-block API
    def encode_blocks(self, data, states):
        """data [nblk][2][N] float64 raw (prior|current) blocks; states: list of [extraBits, bitDeposit] (updated)."""
        data = np.ascontiguousarray(data, dtype=np.float64)
        nblk = data.shape[0]
        NB, M = self.nBands, self.M
        st = (PacStreamState * nblk)()
        for i, s in enumerate(states):
            st[i].extraBits = int(s[0]); st[i].bitDeposit = int(s[1])
        sf = np.zeros((nblk, 2, NB), np.int32); ba = np.zeros((nblk, 2, NB), np.int32)
        mant = np.zeros((nblk, 2, M), np.int32); tid = np.zeros((nblk, 2), np.int32)
        osc = np.zeros((nblk, 2), np.int32); lrms = np.zeros(nblk, np.int32)
        ccap = 4096
        chunk = np.zeros((nblk, 2, ccap), np.uint8); cb = np.zeros((nblk, 2), np.int32)
        self._ck(lib().pac_encode_blocks(self.ctx, _p(data, C.c_double), nblk, st, _p(sf, C.c_int32), _p(ba, C.c_int32),
                                         _p(mant, C.c_int32), _p(tid, C.c_int32), _p(osc, C.c_int32), _p(lrms, C.c_int32),
                                         _p(chunk, C.c_uint8), ccap, _p(cb, C.c_int32)))
        for i, s in enumerate(states):
            s[0] = int(st[i].extraBits); s[1] = int(st[i].bitDeposit)
        chunks = [[chunk[i, ch, :cb[i, ch]].tobytes() for ch in range(2)] for i in range(nblk)]
        return {"sf": sf, "ba": ba, "mant": mant, "tableID": tid, "oscale": osc, "lrms": lrms, "chunks": chunks}

    def decode_blocks(self, sf, ba, mant, oscale, lrms):
        sf = np.ascontiguousarray(sf, np.int32); ba = np.ascontiguousarray(ba, np.int32)
        mant = np.ascontiguousarray(mant, np.int32); oscale = np.ascontiguousarray(oscale, np.int32)
        lrms = np.ascontiguousarray(lrms, np.int32)
        nblk = lrms.shape[0]
        out = np.zeros((nblk, 2, self.N))
        self._ck(lib().pac_decode_blocks(self.ctx, _p(sf, C.c_int32), _p(ba, C.c_int32), _p(mant, C.c_int32),
                                         _p(oscale, C.c_int32), _p(lrms, C.c_int32), nblk, _p(out, C.c_double)))
        return out

    def unpack_blocks(self, chunks):
        """chunks: list of (bytes ch0, bytes ch1)."""
        nblk = len(chunks)
        ccap = max(max(len(c[0]), len(c[1])) for c in chunks) + 8
        buf = np.zeros((nblk, 2, ccap), np.uint8); cb = np.zeros((nblk, 2), np.int32)
        for i, c in enumerate(chunks):
            for ch in range(2):
                buf[i, ch, :len(c[ch])] = np.frombuffer(c[ch], np.uint8); cb[i, ch] = len(c[ch])
        NB, M = self.nBands, self.M
        sf = np.zeros((nblk, 2, NB), np.int32); ba = np.zeros((nblk, 2, NB), np.int32)
        mant = np.zeros((nblk, 2, M), np.int32); osc = np.zeros((nblk, 2), np.int32)
        lrms = np.zeros(nblk, np.int32); tid = np.zeros((nblk, 2), np.int32)
        self._ck(lib().pac_unpack_blocks(self.ctx, _p(buf, C.c_uint8), ccap, _p(cb, C.c_int32), nblk, _p(sf, C.c_int32),
                                         _p(ba, C.c_int32), _p(mant, C.c_int32), _p(osc, C.c_int32), _p(lrms, C.c_int32),
                                         _p(tid, C.c_int32)))
        return {"sf": sf, "ba": ba, "mant": mant, "oscale": osc, "lrms": lrms, "tableID": tid}

    # ---------------------------------------------------------------- L2 entry points
    def window(self, kind, x):
        x = np.ascontiguousarray(x, np.float64)
        x2 = x.reshape(-1, x.shape[-1])
        self._ck(lib().pac_window(self.ctx, kind, _p(x2, C.c_double), x2.shape[0], x2.shape[1]))
        return x

    def mdct(self, x):
        x = np.ascontiguousarray(x, np.float64)
        x2 = x.reshape(-1, x.shape[-1])
        X = np.zeros((x2.shape[0], x2.shape[1] // 2))
        self._ck(lib().pac_mdct(self.ctx, _p(x2, C.c_double), x2.shape[0], x2.shape[1], _p(X, C.c_double)))
        return X.reshape(x.shape[:-1] + (x.shape[-1] // 2,))

    def imdct(self, X):
        X = np.ascontiguousarray(X, np.float64)
        X2 = X.reshape(-1, X.shape[-1])
        x = np.zeros((X2.shape[0], X2.shape[1] * 2))
        self._ck(lib().pac_imdct(self.ctx, _p(X2, C.c_double), X2.shape[0], X2.shape[1] * 2, _p(x, C.c_double)))
        return x.reshape(X.shape[:-1] + (X.shape[-1] * 2,))

    def analysis(self, data):
        """data [nblk][2][N] raw blocks -> dict(lrms, oscale, mdct, bthr, smr, lines)"""
        data = np.ascontiguousarray(data, np.float64)
        nblk = data.shape[0]
        NB, M = self.nBands, self.M
        lrms = np.zeros(nblk, np.int32); osc = np.zeros((nblk, 2), np.int32)
        mdct = np.zeros((nblk, 2, M)); bthr = np.zeros((nblk, 6, M)); smr = np.zeros((nblk, 2, NB)); lines = np.zeros((nblk, 2, M))
        self._ck(lib().pac_analysis(self.ctx, _p(data, C.c_double), nblk, _p(lrms, C.c_int32), _p(osc, C.c_int32),
                                    _p(mdct, C.c_double), _p(bthr, C.c_double), _p(smr, C.c_double), _p(lines, C.c_double)))
        return {"lrms": lrms, "oscale": osc, "mdct": mdct, "bthr": bthr, "smr": smr, "lines": lines}

    def calc_smrs(self, data, mdct, scale):
        data = np.ascontiguousarray(data, np.float64).reshape(-1, self.N)
        mdct = np.ascontiguousarray(mdct, np.float64).reshape(-1, self.M)
        smr = np.zeros((data.shape[0], self.nBands))
        self._ck(lib().pac_calc_smrs(self.ctx, _p(data, C.c_double), _p(mdct, C.c_double), data.shape[0], int(scale), _p(smr, C.c_double)))
        return smr

    def masked_threshold(self, data, noDrop=False):
        data = np.ascontiguousarray(data, np.float64).reshape(-1, self.N)
        thr = np.zeros((data.shape[0], self.M))
        self._ck(lib().pac_masked_threshold(self.ctx, _p(data, C.c_double), data.shape[0], 1 if noDrop else 0, _p(thr, C.c_double)))
        return thr

    def huffman_select(self, mags, ba_per_symbol):
        mags = np.ascontiguousarray(mags, np.uint32); ba = np.ascontiguousarray(ba_per_symbol, np.int32)
        tid = C.c_int32(); tot = np.zeros(NTABLES, np.int64)
        self._ck(lib().pac_huffman_select(self.ctx, _p(mags, C.c_uint32), _p(ba, C.c_int32), len(mags), C.byref(tid), _p(tot, C.c_int64)))
        return int(tid.value), tot

    def bitalloc(self, bitBudget, extraBits, maxMantBits, smr, lrms_mask):
        smr = np.ascontiguousarray(smr, np.float64).reshape(-1, self.nBands)
        n = smr.shape[0]
        bb = np.ascontiguousarray(np.broadcast_to(np.asarray(bitBudget, np.float64), (n,)))
        eb = np.ascontiguousarray(np.broadcast_to(np.asarray(extraBits, np.int64), (n,)))
        lm = np.ascontiguousarray(np.broadcast_to(np.asarray(lrms_mask, np.int32), (n,)))
        bits = np.zeros((n, self.nBands), np.int32); diff = np.zeros(n, np.int64)
        self._ck(lib().pac_bitalloc(self.ctx, n, _p(bb, C.c_double), _p(eb, C.c_int64), int(maxMantBits), _p(smr, C.c_double),
                                    _p(lm, C.c_int32), _p(bits, C.c_int32), _p(diff, C.c_int64)))
        return bits, diff

    def histogram(self, codes, nbins=1 << 16, base=0):
        """Huffman.py:71-83 on the device: (counts[nbins], first[nbins]) of unsigned mantissa codes; `codes` is a uint32 numpy
        array or a CUDA tensor (torch.int32 / uint32 viewed as such).  first = base + index of the first occurrence, -1 if absent."""
        counts = np.zeros(nbins, np.int64); first = np.zeros(nbins, np.int64)
        if hasattr(codes, "data_ptr"):
            ptr, n = C.c_void_p(codes.data_ptr()), int(codes.numel())
        else:
            codes = np.ascontiguousarray(codes, np.uint32)
            ptr, n = codes.ctypes.data_as(C.c_void_p), int(codes.size)
        self._ck(lib().pac_histogram(self.ctx, ptr, n, int(base), int(nbins), _p(counts, C.c_int64), _p(first, C.c_int64)))
        return counts, first

    def bitalloc_alt(self, mode, bitBudget, maxMantBits, level=None):
        """bitalloc.py:22-125 on n problems; mode 'uniform' | 'constsnr' | 'constmnr'; level [n][nBands] (peak SPL / SMR)"""
        m = {"uniform": 0, "constsnr": 1, "constmnr": 2}[mode]
        bb = np.ascontiguousarray(np.atleast_1d(np.asarray(bitBudget, np.float64)))
        n = bb.shape[0]
        lv = None
        if m:
            lv = np.ascontiguousarray(np.broadcast_to(np.asarray(level, np.float64).reshape(-1, self.nBands), (n, self.nBands)))
        bits = np.zeros((n, self.nBands), np.int32)
        self._ck(lib().pac_bitalloc_alt(self.ctx, m, n, _p(bb, C.c_double), int(maxMantBits), _p(lv, C.c_double) if m else None,
                                        _p(bits, C.c_int32)))
        return bits

    def scale_factor(self, x, nScaleBits, nMantBits):
        x = np.ascontiguousarray(np.atleast_1d(x), np.float64)
        out = np.zeros(len(x), np.int32)
        self._ck(lib().pac_scale_factor(self.ctx, _p(x, C.c_double), len(x), int(nScaleBits), int(nMantBits), _p(out, C.c_int32)))
        return out

    def vquantize_uniform(self, x, nBits):
        x = np.ascontiguousarray(x, np.float64)
        q = np.zeros(len(x), np.uint64)
        self._ck(lib().pac_vquantize_uniform(self.ctx, _p(x, C.c_double), len(x), int(nBits), _p(q, C.c_uint64)))
        return q

    def vdequantize_uniform(self, q, nBits):
        q = np.ascontiguousarray(q, np.uint64)
        x = np.zeros(len(q))
        self._ck(lib().pac_vdequantize_uniform(self.ctx, _p(q, C.c_uint64), len(q), int(nBits), _p(x, C.c_double)))
        return x

    def vmantissa(self, x, scale, nScaleBits, nMantBits):
        x = np.ascontiguousarray(x, np.float64)
        m = np.zeros(len(x), np.uint64)
        self._ck(lib().pac_vmantissa(self.ctx, _p(x, C.c_double), len(x), int(scale), int(nScaleBits), int(nMantBits), _p(m, C.c_uint64)))
        return m

    def vdequantize(self, scale, m, nScaleBits, nMantBits):
        m = np.ascontiguousarray(m, np.int64)
        x = np.zeros(len(m))
        self._ck(lib().pac_vdequantize(self.ctx, int(scale), _p(m, C.c_int64), len(m), int(nScaleBits), int(nMantBits), _p(x, C.c_double)))
        return x


# ------------------------------------------------------------------ engine cache for the shim modules
_engines = {}
_default = {"device": int(os.environ.get("PAC_DEVICE", "0")), "precision": os.environ.get("PAC_PRECISION", "fp64")}


def set_default(device=None, precision=None):
    if device is not None:
        _default["device"] = int(device)
    if precision is not None:
        _default["precision"] = precision


def engine(sampleRate=44100, nMDCTLines=1024, nScaleBits=4, nMantSizeBits=4, nTableIDBits=4, targetBitsPerSample=2.27,
           device=None, precision=None, window="sine"):
    key = (device if device is not None else _default["device"], precision or _default["precision"], int(sampleRate),
           int(nMDCTLines), int(nScaleBits), int(nMantSizeBits), int(nTableIDBits), float(targetBitsPerSample), window)
    e = _engines.get(key)
    if e is None:
        e = Engine(key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7], window=window)
        _engines[key] = e
    return e


def engine_for(cp):
    """Engine matching a reference-style CodingParams attribute bag."""
    return engine(getattr(cp, "sampleRate", 44100), getattr(cp, "nMDCTLines", 1024), getattr(cp, "nScaleBits", 4),
                  getattr(cp, "nMantSizeBits", 4), getattr(cp, "nTableIDBits", 4), getattr(cp, "targetBitsPerSample", 2.27),
                  window=getattr(cp, "window", "sine"))
