"""
model_analysis.py -- numpy model of the RESTRUCTURED analysis math the CUDA kernel uses.

Test infrastructure.  The CUDA analysis kernel (csrc/analysis.cuh) does not evaluate the reference's
ten 2048-point FFTs per stereo block; it uses the algebra below.  This model states that algebra in
numpy so that `tests/test_model.py` can prove on the CPU, against dumps of the reference itself
(tests/golden/stages.npz), that the restructuring is exact to rounding:

  * real 2048-pt FFTs as 1024-pt complex FFTs of (even + i*odd) + a split pass,
  * MDCT as fold -> 512-pt complex FFT with pre/post twiddles (DCT-IV), instead of mdct.py:62-71,
  * the Hann^2 / Hann^3 windowed M/S spectra (psychoac.py:549-562 and the in-place window aliasing,
    SURVEY Appendix A Q1) as repeated 3-tap convolutions in the frequency domain of (F_L +- F_R)/2,
  * peak picking on squared magnitudes.
"""
import numpy as np


def rfft_packed(x):
    """X[0..M] of real x (len 2M) from one M-point complex FFT."""
    M = len(x) // 2
    z = x[0::2] + 1j * x[1::2]
    Z = np.fft.fft(z)
    Zx = np.concatenate([Z, Z[:1]])
    k = np.arange(M + 1)
    Zc = np.conj(Zx[M - k])
    E = (Zx[k] + Zc) / 2
    O = (Zx[k] - Zc) / 2j
    return E + np.exp(-2j * np.pi * k / (2 * M)) * O


def mdct_fold_fft(xw):
    """(2/N) * MDCT of the (already windowed) block xw, N = len(xw), via an N/4-point complex FFT."""
    N = len(xw)
    M = N // 2
    H = M // 2
    a, b, c, d = xw[:H], xw[H:2 * H], xw[2 * H:3 * H], xw[3 * H:]
    u = np.concatenate([-c[::-1] - d, a - b[::-1]])
    n = np.arange(H)
    t = (u[2 * n] + 1j * u[M - 1 - 2 * n]) * np.exp(-1j * np.pi * n / M)
    y = np.fft.fft(t) * np.exp(-1j * np.pi * (n + 0.25) / M)
    Y = np.empty(M)
    Y[2 * n] = y.real
    Y[M - 1 - 2 * n] = -y.imag
    return (2.0 / N) * Y


def hann_taps(F, N):
    """DFT of hann(n+1/2)*y from F = DFT(y)[0..N/2] for real y (uses F[-1] = conj(F[1]), F[N/2+1] = conj(F[N/2-1]))."""
    M = N // 2
    Fm1 = np.concatenate([np.conj(F[1:2]), F[:-1]])
    Fp1 = np.concatenate([F[1:], np.conj(F[M - 1:M])])
    w = np.exp(1j * np.pi / N)
    return 0.5 * F - 0.25 * (w * Fm1 + np.conj(w) * Fp1)


def bark(f):
    khz = f / 1000.0
    return 13.0 * np.arctan(khz * 0.76) + 3.5 * np.arctan((khz / 7.5) ** 2)


def thresh(f):
    khz = np.clip(f, 10, np.inf) / 1000.0
    return 3.64 * khz ** -0.8 - 6.5 * np.exp(-0.6 * (khz - 3.3) ** 2) + 0.001 * khz ** 4


def spl(i):
    return np.maximum(96 + 10 * np.log10(np.maximum(i, 10 ** -12.6)), -30.0)


class Tables:
    def __init__(self, N=2048, fs=44100):
        M = N // 2
        f = fs / 2.0 / M * (np.arange(M) + 0.5)
        self.zline = bark(f)
        self.tiq = 10 ** ((thresh(f) - 96) / 10)
        f2 = ((np.arange(M) + 0.5) / M) * (fs / 2.0)
        m = 10.0 ** (1.25 * (1 - np.cos(np.pi * (np.minimum(f2, 3000.) / 3000.)) - 2.5))
        self.mld = m / m.max()
        self.zpeak = bark(np.arange(M) * float(fs // N))
        self.sine = np.sin((np.arange(N) + 0.5) * np.pi / N)
        self.hann = 0.5 * (1 - np.cos(2.0 * (np.arange(N) + 0.5) * np.pi / N))
        self.N, self.M, self.fs = N, M, fs


def curve(T, F, drop):
    """masked threshold (dB) at the MDCT lines from spectrum F[0..M] (psychoac.py:431-456)."""
    M, N = T.M, T.N
    P = (F.real ** 2 + F.imag ** 2)[:M]
    k = np.arange(1, M - 1)
    pk = k[(P[k] > P[k - 1]) & (P[k] > P[k + 1]) & (P[k] > 1e-6)]
    acc = np.zeros(M)
    cs = np.concatenate([[0.0], np.cumsum(P)])
    for kk in pk:
        s = 0.0 if kk < 3 else np.sum(P[kk - 3:min(kk + 3, M)])
        Pm = float(spl(np.array(8.0 / 3.0 * 4.0 / N ** 2 * s)))
        dz = T.zline - T.zpeak[kk]
        lev = 0.367 * max(Pm - 40.0, 0)
        spread = (np.where(dz >= 0, lev, 0.0) - 27.0) * np.maximum(np.abs(dz) - 0.5, 0.0)
        acc += 10 ** ((Pm + spread - drop - 96) / 10)
    return spl(acc + T.tiq)


def analysis(T, xl, xr, nLines):
    """raw signed-fraction blocks -> (lrms, oscale, mdct_scaled[2], bthr6, smr[2][nb], lines[2][M])"""
    N, M = T.N, T.M
    lo = np.concatenate([[0], np.cumsum(nLines)[:-1]])
    # codec.py:96-102
    L, R = rfft_packed(xl)[:M], rfft_packed(xr)[:M]
    lrms = np.zeros(len(nLines), dtype=int)
    for b, (l0, n) in enumerate(zip(lo, nLines)):
        d = np.sum(L[l0:l0 + n] ** 2 - R[l0:l0 + n] ** 2)
        s = np.sum(L[l0:l0 + n] ** 2 + R[l0:l0 + n] ** 2)
        lrms[b] = abs(d) < 0.8 * abs(s)
    xs = [xl * T.sine, xr * T.sine]
    X, osc = [], []
    for ch in range(2):
        Xc = mdct_fold_fft(xs[ch])
        q = min(int(((2 ** 20 - 1) * min(np.max(np.abs(Xc)), 1.0) + 1) / 2), 2 ** 19 - 1)
        s = 15 if q == 0 else min(15, 18 - (q.bit_length() - 1))
        osc.append(s)
        X.append(Xc * (1 << s))
    F1 = [rfft_packed(xs[ch] * T.hann) for ch in range(2)]
    F2 = [hann_taps((F1[0] + F1[1]) / 2, N), hann_taps((F1[0] - F1[1]) / 2, N)]
    F3 = [hann_taps(F2[0], N), hann_taps(F2[1], N)]
    bthr = np.stack([curve(T, F1[0], 15.0), curve(T, F1[1], 15.0), curve(T, F2[0], 15.0),
                     curve(T, F2[1], 15.0), curve(T, F3[0], 0.0), curve(T, F3[1], 0.0)])
    XM, XS = (X[0] + X[1]) / 2, (X[0] - X[1]) / 2
    sp = lambda x, s: spl(4.0 * x * x) - 6.02 * s
    thrM = np.maximum(bthr[2], np.minimum(bthr[3], bthr[5] * T.mld))
    thrS = np.maximum(bthr[3], np.minimum(bthr[2], bthr[4] * T.mld))
    vLR = [sp(X[0], osc[0]) - bthr[0], sp(X[1], osc[1]) - bthr[1]]
    vMS = [sp(XM, osc[0]) - thrM, sp(XS, osc[1]) - thrS]
    smr = np.zeros((2, len(nLines)))
    lines = np.zeros((2, M))
    for b, (l0, n) in enumerate(zip(lo, nLines)):
        for ch in range(2):
            v = (vMS if lrms[b] else vLR)[ch][l0:l0 + n]
            smr[ch, b] = np.max(v) if n else -96.0
            lines[ch, l0:l0 + n] = ((XM, XS) if lrms[b] else X)[ch][l0:l0 + n]
    return lrms, np.array(osc), np.stack(X), bthr, smr, lines


# ---------------------------------------------------------------------------------------------------------------
# Scan-based threshold evaluation used by the fp32 fast mode (csrc/analysis.cuh: masked_curve_fast).
#
# The reference sums, for every masker, an exp over all 1024 lines (O(maskers x lines)).  The spreading function is
# piecewise: plateau (|dz| <= .5 Bark), lower skirt (fixed -27 dB/Bark) and upper skirt (-27 + .367 max(P-40, 0)
# dB/Bark).  Fixed-slope parts are separable, so they become weighted scans with STATIC weights (curve_v3 below: over the
# bins); only the upper skirts of maskers louder than 40 dB keep a pairwise evaluation, over the lines above them, and
# culled per 64-line half-chunk.
# ---------------------------------------------------------------------------------------------------------------

K10 = np.log2(10.0) / 10.0


class Geometry:
    """static line/bin geometry of the v2 evaluation (all from the Bark tables)"""

    def __init__(self, T):
        M = T.M
        zl, zp = T.zline, T.zpeak
        # eL[k]: highest line strictly inside bin k's lower skirt (z_i < zp_k - .5), -1 if none
        # eU[k]: lowest line strictly inside bin k's upper skirt (z_i > zp_k + .5), M if none
        self.eL = np.searchsorted(zl, zp - 0.5, side="left") - 1          # z_i <  zp-.5  <=> i <= eL
        self.eU = np.searchsorted(zl, zp + 0.5, side="right")             # z_i >  zp+.5  <=> i >= eU
        # careful with the reference's comparisons: plateau is |dz| <= .5 exactly (psychoac.py:116 uses > .5)
        dz = zl[None, :] - zp[:, None]
        lower = dz < -0.5
        upper = dz > 0.5
        assert all((np.nonzero(lower[k])[0].max(initial=-1) == self.eL[k]) for k in range(M))
        assert all((np.nonzero(upper[k])[0].min(initial=M) == self.eU[k]) for k in range(M))
        self.dn = -27.0 * K10
        # gap from the skirt's origin to its entry line, as exponent factors
        self.gL = np.where(self.eL >= 0, zp - 0.5 - zl[np.maximum(self.eL, 0)], 0.0)      # >= 0
        self.gU = np.where(self.eU < M, zl[np.minimum(self.eU, M - 1)] - zp - 0.5, 0.0)   # >= 0
        # for every line: range of bins whose entry line it is (monotone maps -> contiguous ranges)
        self.kLa = np.searchsorted(self.eL, np.arange(M), side="left")
        self.kLb = np.searchsorted(self.eL, np.arange(M), side="right")
        self.kUa = np.searchsorted(self.eU, np.arange(M), side="left")
        self.kUb = np.searchsorted(self.eU, np.arange(M), side="right")
        # plateau of line i: bins k with eL[k] < i < eU[k]  <=> k in [pa_i, pb_i)
        self.pa = np.searchsorted(self.eU, np.arange(M), side="right")    # first k with eU[k] > i
        self.pb = np.searchsorted(self.eL, np.arange(M), side="left")     # first k with eL[k] >= i
        self.M = M


def weighted_suffix_scan(U, z, dn, dtype):
    """L[i] = sum_{j>=i} U[j] 2^{dn (z_j - z_i)} evaluated the way the kernel does: 4 entries per thread, warp
    Kogge-Stone with static weights, block carry (z = the Bark positions of the scan's domain: the bins for curve_v3)."""
    M = len(U)
    w = lambda i, j: dtype(2.0 ** (dn * (z[j] - z[i])))
    a = np.zeros(M, dtype)
    nthr = M // 4
    g = np.zeros(nthr, dtype)
    for t in range(nthr):
        b = 4 * t
        a[b + 3] = U[b + 3]
        for q in (2, 1, 0):
            a[b + q] = dtype(U[b + q] + a[b + q + 1] * w(b + q, b + q + 1))
        g[t] = a[b]
    nw = nthr // 32
    for wi in range(nw):
        base = wi * 32
        for s in (1, 2, 4, 8, 16):
            old = g.copy()
            for l in range(32):
                if l + s < 32:
                    t = base + l
                    g[t] = dtype(old[t] + old[t + s] * w(4 * t, 4 * (t + s)))
    tot = np.array([g[wi * 32] for wi in range(nw)], dtype)
    out = np.zeros(M, dtype)
    for wi in range(nw):
        C = dtype(0)
        for w2 in range(nw - 1, wi, -1):
            nxt = 128 * (w2 + 1)
            C = dtype(tot[w2] + C * (w(128 * w2, nxt) if nxt < M else dtype(0)))
        for l in range(32):
            t = wi * 32 + l
            nb = 4 * (t + 1)
            inc = dtype(g[t + 1]) if l < 31 else dtype(0)
            if wi < nw - 1:
                inc = dtype(inc + C * w(nb, 128 * (wi + 1))) if nb <= 128 * (wi + 1) else inc
            for q in range(4):
                i = 4 * t + q
                out[i] = dtype(a[i] + (inc * w(i, nb) if nb < M else dtype(0)))
    return out


# ---------------------------------------------------------------------------------------------------------------
# curve_v3: the fixed-slope skirts as scans over the BINS (round 1 scanned over the lines, gathering up to three bins per line).  A masker at bin k spreads downwards as
# A_k 2^{dn (zp_k - .5 - z_i)} over the lines i <= eL[k].  With pb(i) = first bin whose lower skirt reaches line i (the same index
# that bounds the plateau window [pa, pb)), the sum over maskers factorises:
#     low(i) = SD[pb(i)] * 2^{dn (zp_pb - .5 - z_i)},   SD[k] = sum_{k' >= k} A_k' 2^{dn (zp_k' - zp_k)}   (a scan over bins)
# and likewise the quiet upper skirts, up(i) = SA[pa(i) - 1] * 2^{dn (z_i - zp_{pa-1} - .5)}, SA an ascending scan over the quiet
# maskers.  The scans run on the values the peak-picking thread already holds (no dense injection arrays, no per-line gathers of
# up to three bins, no per-masker entry exponentials); a line needs one entry of each scan and a static factor <= 1.
# ---------------------------------------------------------------------------------------------------------------

def curve_v3(T, G, F, drop, dtype=np.float32, stats=None):
    M, N = T.M, T.N
    P = (F.real.astype(dtype) ** 2 + F.imag.astype(dtype) ** 2)[:M]
    k = np.arange(1, M - 1)
    pk = k[(P[k] > P[k - 1]) & (P[k] > P[k + 1]) & (P[k] > dtype(1e-6))]
    A = np.zeros(M, dtype); Aq = np.zeros(M, dtype)
    loud = []
    cn = dtype(8.0 / 3.0 * 4.0 / N ** 2)
    for kk in pk:
        s = dtype(0) if kk < 3 else np.sum(P[kk - 3:min(kk + 3, M)], dtype=dtype)
        Pm = dtype(max(96 + 10 * np.log10(max(dtype(cn * s), dtype(10 ** -12.6))), -30.0))
        c0 = dtype((Pm - dtype(drop) - dtype(96)) * dtype(K10))
        A[kk] = dtype(2.0 ** c0)
        lev = dtype(0.367) * max(Pm - dtype(40), dtype(0))
        if lev > 0:
            if G.eU[kk] < M:
                loud.append((kk, c0, dtype((lev - dtype(27)) * dtype(K10))))
        else:
            Aq[kk] = A[kk]
    zp, zl = T.zpeak, T.zline
    SD = weighted_suffix_scan(A, zp, G.dn, dtype)                                  # SD[k] = sum_{k' >= k} A 2^{dn (zp_k' - zp_k)}
    SA = weighted_suffix_scan(Aq[::-1].copy(), -zp[::-1], G.dn, dtype)[::-1]        # SA[k] = sum_{k' <= k} Aq 2^{dn (zp_k - zp_k')}
    SDx = np.concatenate([SD, [dtype(0)]])                                          # SD[M] = 0
    SAx = np.concatenate([[dtype(0)], SA])                                          # SA[-1] = 0 (index shifted by one)
    pa, pb = G.pa, G.pb
    fL = np.where(pb < M, 2.0 ** (G.dn * (zp[np.minimum(pb, M - 1)] - 0.5 - zl)), 0.0).astype(dtype)
    fU = np.where(pa > 0, 2.0 ** (G.dn * (zl - zp[np.maximum(pa - 1, 0)] - 0.5)), 0.0).astype(dtype)
    assert fL.max() <= 1.0 and fU.max() <= 1.0
    Sp = np.concatenate([[0.0], np.cumsum(A.astype(np.float64))])                   # double prefix (plateau = exact range sums)
    plat = (Sp[pb] - Sp[pa]).astype(dtype)
    acc = (((SDx[pb] * fL).astype(dtype) + (SAx[pa] * fU).astype(dtype)).astype(dtype) + plat).astype(dtype)
    acc = (acc + T.tiq.astype(dtype)).astype(dtype)
    npairs = nculled = 0
    part = acc.copy()
    for kk, c0, up in loud:
        i0 = G.eU[kk]
        B = np.float64(c0) - 0.5 * np.float64(up) - np.float64(up) * T.zpeak[kk]
        for h in range(M // 64):
            lo, hi = 64 * h, 64 * h + 64
            if i0 >= hi:
                continue
            cut = np.log2(np.float64(part[lo:hi].min())) - 26.01
            if dtype(np.float64(up) * dtype(zl[lo]) + dtype(B)) < cut:
                nculled += hi - max(lo, i0)
                continue
            a = max(lo, i0)
            e = (np.float64(up) * zl[a:hi] + B).astype(dtype)
            acc[a:hi] += (2.0 ** e.astype(np.float64)).astype(dtype)
            npairs += hi - a
    if stats is not None:
        stats["peaks"] = stats.get("peaks", 0) + len(pk)
        stats["loud"] = stats.get("loud", 0) + len(loud)
        stats["pairs"] = stats.get("pairs", 0) + npairs
        stats["culled"] = stats.get("culled", 0) + nculled
    return np.maximum(dtype(96) + dtype(10) * np.log10(np.maximum(acc, dtype(10 ** -12.6))).astype(dtype), dtype(-30))
