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
