"""corpus.py -- the synthetic corpus of BASELINE config 3/4, reproducible bit for bit on CPU and GPU.

The reference has no generator; the survey defines one: stream `s` is keyed with 0x50414300 + s ('PAC\\0' + stream id) and is the sum of
3-8 sinusoids (log-uniform 50 Hz..16 kHz, -30..-6 dBFS, independent L/R gains), noise at -50..-25 dBFS (per-channel fraction random) and
Poisson(2/s) transients (5 ms exponentially decaying noise bursts at -12..-3 dBFS), clipped and rounded to int16.

Two layers, so that "the same inputs" does not rest on which device drew them (round-1 review, weak item 9):
  * per-stream PARAMETERS (a few dozen scalars) come from NumPy's `Generator(Philox(key))` on the host, whatever the device;
  * per-sample values are pure INTEGER arithmetic on a counter-based Philox4x32-10 (key = (stream key, domain), counter = sample
    index): the same int64 expressions run under numpy and under torch (CPU or CUDA), so every backend produces identical int16
    samples -- no transcendental is evaluated per sample.  Sines come from a 4096-entry table (built once on the host in float64,
    rounded to Q24) with linear interpolation of a 32-bit phase accumulator; the noise is the sum of eight uniform bytes per sample
    and channel (Irwin-Hall, sigma = 209.0 units: a bell-shaped approximation of Gaussian noise, tails end at +-4.9 sigma).

`gen_streams(ids, n, device)` -> torch int16 [len(ids)][n][2] on `device`; `gen_streams_numpy(ids, n)` -> the same as a numpy array.
tests/test_host.py::test_corpus_generator_is_backend_independent compares the two (and the CUDA backend when a GPU is present).
"""
import numpy as np

FS = 44100
KEY0 = 0x50414300
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF
_TAB_BITS = 12
_NOISE_SIGMA = float(np.sqrt(8 * (256.0 ** 2 - 1) / 12.0))      # of the sum of eight uniform bytes
_BLEN = int(0.005 * FS)


def _mulhilo(a, b):
    """(hi, lo) 32-bit halves of a * b for a 32-bit constant `a` and int64 array `b` holding 32-bit values -- without ever leaving
    the positive int64 range (torch has no uint64 arithmetic)."""
    bl = b & 0xFFFF
    bh = b >> 16
    pl = bl * a                      # < 2^48
    ph = bh * a                      # < 2^48
    hi = (ph + (pl >> 16)) >> 16
    lo = (pl + ((ph & 0xFFFF) << 16)) & _MASK
    return hi, lo


def philox4x32(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11).  Counters/keys: int64 arrays (or Python ints) holding 32-bit values; numpy or torch."""
    for _ in range(10):
        hi0, lo0 = _mulhilo(_M0, c0)
        hi1, lo1 = _mulhilo(_M1, c2)
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


def _byte_sums(words):
    """Sum of the eight bytes of two 32-bit words, minus the mean (1020): the noise sample of one channel."""
    a, b = words
    s = (a & 0xFF) + ((a >> 8) & 0xFF) + ((a >> 16) & 0xFF) + (a >> 24)
    s = s + (b & 0xFF) + ((b >> 8) & 0xFF) + ((b >> 16) & 0xFF) + (b >> 24)
    return s - 1020


_SINE_TABLE = None


def sine_table():
    global _SINE_TABLE
    if _SINE_TABLE is None:
        j = np.arange((1 << _TAB_BITS) + 1, dtype=np.float64)
        _SINE_TABLE = np.round(np.sin(2.0 * np.pi * j / (1 << _TAB_BITS)) * (1 << 24)).astype(np.int64)
    return _SINE_TABLE


def stream_params(s, n):
    """Host-side draw of one stream's parameters (all integers after this point)."""
    rng = np.random.Generator(np.random.Philox(key=KEY0 + int(s)))
    ntones = int(rng.integers(3, 9))
    f = 50.0 * (320.0 ** rng.random(ntones))
    amp = 10.0 ** (-(6.0 + 24.0 * rng.random((ntones, 2))) / 20.0)
    ph = rng.random(ntones)
    level = 10.0 ** (-(25.0 + 25.0 * rng.random()) / 20.0)
    frac = rng.random(2)
    nb = int(rng.poisson(2.0 * n / FS))
    if n <= _BLEN:
        nb = 0
    pos = (rng.random(nb) * max(n - _BLEN, 1)).astype(np.int64)
    lvl = 10.0 ** (-(3.0 + 9.0 * rng.random(nb)) / 20.0)
    return {
        "dphi": np.round(f / FS * 2.0 ** 32).astype(np.int64) & _MASK,
        "ph0": np.round(ph * 2.0 ** 32).astype(np.int64) & _MASK,
        "ampq": np.round(amp * 65536.0).astype(np.int64),                                   # Q16 gains per (tone, channel)
        "noiseq": np.round(level * frac * (1 << 24) / _NOISE_SIGMA).astype(np.int64),       # per channel: Q24 units per noise unit
        "pos": pos,
        "lvlq": np.round(lvl * (1 << 24) / _NOISE_SIGMA).astype(np.int64),                  # per burst
    }


def _decay_q15():
    return np.round(np.exp(-np.arange(_BLEN, dtype=np.float64) / (_BLEN / 4.0)) * 32768.0).astype(np.int64)


def _synth(s, n, xp, dev):
    """One stream as int64 Q0 sample values in [-32767, 32767], shape [n][2], in backend `xp` ('np' or torch module)."""
    p = stream_params(s, n)
    key = (KEY0 + int(s)) & _MASK
    if xp is np:
        arange = lambda m: np.arange(m, dtype=np.int64)
        asarr = lambda a: np.asarray(a, dtype=np.int64)
        stack = lambda xs: np.stack(xs, axis=-1)
        clip = np.clip
        take = lambda tab, idx: tab[idx]
    else:
        arange = lambda m: xp.arange(m, dtype=xp.int64, device=dev)
        asarr = lambda a: xp.as_tensor(np.asarray(a, dtype=np.int64), device=dev)
        stack = lambda xs: xp.stack(xs, dim=-1)
        clip = lambda a, lo, hi: xp.clamp(a, lo, hi)
        take = lambda tab, idx: tab[idx]
    t = arange(n)
    tab = asarr(sine_table())
    acc = [None, None]
    # tones: 32-bit phase accumulator -> table + linear interpolation (Q24), gains in Q16
    for k in range(len(p["dphi"])):
        phase = (t * int(p["dphi"][k]) + int(p["ph0"][k])) & _MASK
        idx = phase >> (32 - _TAB_BITS)
        fr = phase & ((1 << (32 - _TAB_BITS)) - 1)
        lo = take(tab, idx)
        val = lo + (((take(tab, idx + 1) - lo) * fr) >> (32 - _TAB_BITS))
        for ch in range(2):
            term = (val * int(p["ampq"][k, ch])) >> 16
            acc[ch] = term if acc[ch] is None else acc[ch] + term
    # noise: Philox(counter = sample index, key = (stream, 0)) -> 16 bytes -> eight per channel
    w = philox4x32(t, 0, 0, 0, key, 0)
    for ch in range(2):
        acc[ch] = acc[ch] + _byte_sums((w[2 * ch], w[2 * ch + 1])) * int(p["noiseq"][ch])
    sig = stack(acc)                                                            # [n][2] Q24
    q = (clip(sig, -(1 << 24), 1 << 24) * 32767 + (1 << 23)) >> 24              # round(clamp(x, -1, 1) * 32767), half up
    # transients: Philox(counter = (sample in burst, burst index), key = (stream, 1)); rounded to integers before they are added
    nb = len(p["pos"])
    if nb:
        j = arange(_BLEN)
        dec = asarr(_decay_q15())
        bi = arange(nb)
        c0 = j[None, :] + 0 * bi[:, None]
        c1 = bi[:, None] + 0 * j[None, :]
        w = philox4x32(c0, c1, 0, 0, key, 1)
        lvl = asarr(p["lvlq"])
        pos = asarr(p["pos"])
        bursts = []
        for ch in range(2):
            g = _byte_sums((w[2 * ch], w[2 * ch + 1]))                          # [nb][BLEN]
            b24 = (g * lvl[:, None] * dec[None, :]) >> 15                       # Q24
            bursts.append((clip(b24, -(1 << 24), 1 << 24) * 32767 + (1 << 23)) >> 24)
        burst = stack(bursts).reshape(-1, 2)
        idx = (pos[:, None] + j[None, :]).reshape(-1)
        if xp is np:
            np.add.at(q, idx, burst)
        else:
            q.index_add_(0, idx, burst)                                         # integer adds commute: order of arrival does not matter
    return clip(q, -32767, 32767)


def gen_streams_numpy(stream_ids, n):
    out = np.empty((len(stream_ids), n, 2), dtype=np.int16)
    for j, s in enumerate(stream_ids):
        out[j] = _synth(s, n, np, None).astype(np.int16)
    return out


def gen_streams(stream_ids, n, device):
    """int16 [len(ids)][n][2] on `device` (torch)."""
    import torch
    out = torch.empty(len(stream_ids), n, 2, dtype=torch.int16, device=device)
    for j, s in enumerate(stream_ids):
        out[j] = _synth(s, n, torch, device).to(torch.int16)
    return out
