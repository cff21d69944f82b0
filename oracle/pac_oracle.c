/*
 * pac_oracle.c -- CPU restatement of wisamreid/Perceptual-Audio-Codec's per-block hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see pac_oracle.h).  Plain C99 + libm, IEEE double throughout,
 * written to follow the reference's operation order line by line.  Citations are
 * `file:line` relative to /root/reference/codec/.
 *
 * Parity status: PINNED.  tests/test_oracle_*.py check this file against
 *   - the reference's committed whole-file goldens (coded/<n>.wak, outputs/<n>.wav) through
 *     tests/golden/manifest.json + the committed piano_test2 / castanets fixtures,
 *   - per-stage dumps of the (mechanically patched, see oracle/ref_py3.py) reference itself
 *     (tests/golden/stages.npz), and
 *   - the reference's own self-test vectors (tests/golden/kats.json).
 */
#include "pac_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ band layout */

/* psychoac.py:122 */
static const double cbFreqLimits[25] = {100.0, 200.0, 300.0, 400.0, 510.0, 630.0, 770.0, 920.0, 1080.0,
    1270.0, 1480.0, 1720.0, 2000.0, 2320.0, 2700.0, 3150.0, 3700.0, 4400.0, 5300.0, 6400.0, 7700.0,
    9500.0, 12000.0, 15500.0, 24000.0};

/* psychoac.py:124-156.  `sampleRate / 2` at :133 is Python-2 integer division when sampleRate is
 * an int (it is: pcmfile.py:44-45), `sampleRate/2.0` at :141,143 is float. */
int orc_band_layout(int nMDCTLines, int sampleRate, int32_t *nLines)
{
    double half_int = (double)(sampleRate / 2);
    double lower = 0.0;
    for (int b = 0; b < 25; b++) {
        double upper = cbFreqLimits[b] >= sampleRate / 2.0 ? sampleRate / 2.0 : cbFreqLimits[b];
        int cnt = 0;
        for (int i = 0; i < nMDCTLines; i++) {
            double f = (i + 0.5) / nMDCTLines * half_int;
            if (f <= upper && f > lower) cnt++;
        }
        nLines[b] = cnt;
        lower = upper;
    }
    return 25;
}

/* ------------------------------------------------------------------ windows */

/* window.py:27-39 */
void orc_sine_window(double *x, int N)
{
    double Nf = (double)N;
    for (int n = 0; n < N; n++) x[n] *= sin((n + 0.5) * M_PI / Nf);
}

/* window.py:41-53 */
void orc_hann_window(double *x, int N)
{
    double Nf = (double)N;
    for (int n = 0; n < N; n++) x[n] *= 0.5 * (1 - cos(2.0 * (n + 0.5) * M_PI / Nf));
}

/* modified Bessel I0 by its power series (np.i0 uses Chebyshev fits; agreement ~1e-15 rel) */
static double bessel_i0(double x)
{
    double s = 1.0, t = 1.0, q = x * x / 4.0;
    for (int k = 1; k < 500; k++) {
        t *= q / ((double)k * (double)k);
        s += t;
        if (t < s * 1e-18) break;
    }
    return s;
}

/* window.py:56-78 */
void orc_kbd_window(const double *x, double *out, int N, double alpha)
{
    int half = N / 2;
    double *kaiser = (double *)malloc(sizeof(double) * (half + 1));
    double denom = 0.0, den0 = bessel_i0(M_PI * alpha);
    for (int t = 0; t <= half; t++) {
        double u = 4.0 * t / (double)N - 1.0;
        double arg = 1.0 - u * u;
        if (arg < 0) arg = 0;
        kaiser[t] = bessel_i0(alpha * M_PI * sqrt(arg)) / den0;
        denom += kaiser[t] * kaiser[t];
    }
    double c = 0.0;
    for (int t = 0; t < half; t++) {
        c += kaiser[t] * kaiser[t];
        double w = sqrt(c / denom);
        out[t] = x[t] * w;
        out[N - 1 - t] = x[N - 1 - t] * w;
    }
    free(kaiser);
}

/* ------------------------------------------------------------------ FFT (stands in for np.fft) */

void orc_fft(double *re, double *im, int N, int inverse)
{
    /* bit reversal */
    for (int i = 1, j = 0; i < N; i++) {
        int bit = N >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int len = 2; len <= N; len <<= 1) {
        int half = len >> 1;
        for (int k = 0; k < half; k++) {
            double ang = (inverse ? 2.0 : -2.0) * M_PI * k / len;
            double wr = cos(ang), wi = sin(ang);
            for (int i = k; i < N; i += len) {
                int j = i + half;
                double xr = re[j] * wr - im[j] * wi;
                double xi = re[j] * wi + im[j] * wr;
                re[j] = re[i] - xr; im[j] = im[i] - xi;
                re[i] += xr; im[i] += xi;
            }
        }
    }
    if (inverse) {
        double s = 1.0 / N;
        for (int i = 0; i < N; i++) { re[i] *= s; im[i] *= s; }
    }
}

/* ------------------------------------------------------------------ MDCT */

/* mdct.py:49-71 (forward branch) */
void orc_mdct(const double *x, int a, int b, double *X)
{
    int N = a + b;
    double n0 = (b + 1) / 2.;
    double *re = (double *)malloc(sizeof(double) * N * 2), *im = re + N;
    for (int n = 0; n < N; n++) {
        double ang = -2. * M_PI * n / (2. * N);      /* :66 */
        re[n] = x[n] * cos(ang);
        im[n] = x[n] * sin(ang);
    }
    orc_fft(re, im, N, 0);                           /* :68 */
    for (int k = 0; k < N / 2; k++) {
        double ang = (-2. * M_PI / N) * n0 * (k + (1 / 2.));   /* :70 */
        X[k] = (2. / N) * (re[k] * cos(ang) - im[k] * sin(ang));
    }
    free(re);
}

/* mdct.py:73-80 (inverse branch) */
void orc_imdct(const double *X, int a, int b, double *x)
{
    int N = a + b, h = N / 2;
    double n0 = (b + 1) / 2.;
    double *re = (double *)malloc(sizeof(double) * N * 2), *im = re + N;
    for (int k = 0; k < N; k++) {
        double v = k < h ? X[k] : -X[N - 1 - k];     /* hstack((data,-data[::-1])) :75 */
        double ang = 2. * M_PI * k * n0 / N;
        re[k] = v * cos(ang);
        im[k] = v * sin(ang);
    }
    orc_fft(re, im, N, 1);                           /* :77 */
    for (int n = 0; n < N; n++) {
        double ang = 2 * M_PI / (2. * N) * (n + n0); /* :79 */
        x[n] = N * (re[n] * cos(ang) - im[n] * sin(ang));
    }
    free(re);
}

/* ------------------------------------------------------------------ psychoacoustic scalars */

/* psychoac.py:37-42 */
double orc_intensity(double spl) { return pow(10.0, (spl - 96) / 10); }

/* psychoac.py:15-35 */
double orc_spl(double intensity)
{
    double minval = orc_intensity(-30);
    if (intensity < minval) intensity = minval;
    double spl = 96 + 10 * log10(intensity);
    if (spl < -30.) spl = -30.;
    return spl;
}

/* psychoac.py:44-54 */
double orc_thresh(double f)
{
    if (f < 10) f = 10;
    double khz = f / 1000.0;
    double term1 = 3.64 * pow(khz, -0.8);
    double term2 = -6.5 * exp(-0.6 * ((khz - 3.3) * (khz - 3.3)));
    double term3 = 0.001 * pow(khz, 4);
    return term1 + term2 + term3;
}

/* psychoac.py:56-64 */
double orc_bark(double f)
{
    double khz = f / 1000.0;
    double t = khz / 7.5;
    return 13.0 * atan(khz * 0.76) + 3.5 * atan(t * t);
}

/* ------------------------------------------------------------------ quantiser */

/* quantize.py:40-64 */
int orc_quantize_uniform(double x, int nBits)
{
    if (nBits <= 0) return 0;
    int64_t signBitMask = (int64_t)1 << (nBits - 1);
    double largestVal = (double)(signBitMask << 1) - 1.0;
    int64_t q;
    if (fabs(x) >= 1) q = signBitMask - 1;
    else q = (int64_t)((largestVal * fabs(x) + 1.0) / 2.0);
    if (x < 0) q += signBitMask;
    return (int)q;
}

/* quantize.py:91-117 */
void orc_vquantize_uniform(const double *x, int n, int nBits, uint64_t *q)
{
    uint64_t signBitMask = (uint64_t)1 << (nBits - 1);
    double largestVal = (double)(signBitMask << 1) - 1.0;
    for (int i = 0; i < n; i++) {
        double a = fabs(x[i]);
        uint64_t v = a < 1 ? (uint64_t)((a * largestVal + 1.0) / 2.0) : signBitMask - 1;
        if (signbit(x[i])) v += signBitMask;
        q[i] = v;
    }
}

/* quantize.py:120-145 */
void orc_vdequantize_uniform(const uint64_t *q, int n, int nBits, double *x)
{
    uint64_t signBitMask = (uint64_t)1 << (nBits - 1);
    double largestVal = (double)(signBitMask << 1) - 1.0;
    for (int i = 0; i < n; i++) {
        uint64_t v = q[i];
        int neg = (v & signBitMask) == signBitMask;
        if (neg) v -= signBitMask;
        double a = 2.0 * (double)v / largestVal;
        x[i] = neg ? -a : a;
    }
}

/* quantize.py:148-177 */
int orc_scale_factor(double x, int nScaleBits, int nMantBits)
{
    if (nScaleBits < 0) nScaleBits = 0;
    if (nMantBits <= 0) return 0;
    int scale = 0;
    int largestScale = (1 << nScaleBits) - 1;
    int R = nMantBits + largestScale;
    int64_t zeroBitMask = (int64_t)1 << (R - 1);
    int64_t q = (int64_t)orc_quantize_uniform(fabs(x), R) << 1;
    while (scale < largestScale && (zeroBitMask & q) == 0) { q <<= 1; scale++; }
    return scale;
}

/* quantize.py:315-342 */
void orc_vmantissa(const double *x, int n, int scale, int nScaleBits, int nMantBits, uint64_t *m)
{
    if (nScaleBits < 0) nScaleBits = 0;
    uint64_t signBitMask = (uint64_t)1 << (nMantBits - 1);
    int largestScale = (1 << nScaleBits) - 1;
    int R = nMantBits + largestScale;
    uint64_t qmask = (uint64_t)1 << (R - 1);
    double largestVal = (double)(qmask << 1) - 1.0;
    for (int i = 0; i < n; i++) {
        double a = fabs(x[i]);                                  /* :333-335 */
        uint64_t q = a < 1 ? (uint64_t)((a * largestVal + 1.0) / 2.0) : qmask - 1;  /* :337 -> :110-112 */
        uint64_t v = (q << (scale + 1)) >> (R - nMantBits + 1); /* :337-338 */
        if (signbit(x[i])) v += signBitMask;                    /* :340 */
        m[i] = v;
    }
}

/* quantize.py:345-376 */
void orc_vdequantize(int scale, const int64_t *mant, int n, int nScaleBits, int nMantBits, double *x)
{
    if (nScaleBits < 0) nScaleBits = 0;
    int64_t signBitMask = (int64_t)1 << (nMantBits - 1);
    int largestScale = (1 << nScaleBits) - 1;
    int R = nMantBits + largestScale;
    double largestVal = (double)((int64_t)1 << R) - 1.0;
    for (int i = 0; i < n; i++) {
        int64_t m = mant[i];
        int neg = (m & signBitMask) == signBitMask;
        if (neg) m -= signBitMask;
        int64_t q = m << (largestScale - scale);
        if (scale < largestScale && m > 0) q += (int64_t)1 << (largestScale - scale - 1);
        double a = 2.0 * (double)q / largestVal;                /* :374 -> :141 */
        x[i] = neg ? -a : a;
    }
}

/* ------------------------------------------------------------------ bit allocation */

/* bitalloc.py:129-184 */
int orc_bitalloc(double bitBudget, int64_t extraBits, int maxMantBits, int nBands, const int32_t *nLines,
                 const double *SMR, const int32_t *LRMS, int32_t *bits, int64_t *bitDifference)
{
    int valid[ORC_MAX_BANDS];
    int nvalid = nBands;
    for (int b = 0; b < nBands; b++) { bits[b] = 0; valid[b] = 1; }
    int64_t totalBits = (int64_t)(bitBudget + (double)extraBits);   /* int() truncates toward zero */
    while (nvalid > 0) {
        int iMax = -1;
        double best = 0;
        for (int b = 0; b < nBands; b++) {
            if (!valid[b]) continue;
            double v = SMR[b] - bits[b] * 6.;
            if (iMax < 0 || v > best) { best = v; iMax = b; }
        }
        double mx = SMR[0] - (bits[0] - 1) * 6.;
        for (int b = 1; b < nBands; b++) {
            double v = SMR[b] - (bits[b] - 1) * 6.;
            if (v > mx) mx = v;
        }
        if (LRMS[iMax]) { if (mx < -5.0 && valid[iMax]) { valid[iMax] = 0; nvalid--; } }
        else            { if (mx < -15.0 && valid[iMax]) { valid[iMax] = 0; nvalid--; } }
        if (totalBits - nLines[iMax] >= 0) {
            bits[iMax] += 1;
            totalBits -= nLines[iMax];
            if (bits[iMax] >= maxMantBits && valid[iMax]) { valid[iMax] = 0; nvalid--; }
        } else if (valid[iMax]) { valid[iMax] = 0; nvalid--; }
    }
    for (int b = 0; b < nBands; b++)
        if (bits[b] == 1) { totalBits += nLines[b]; bits[b] = 0; }
    *bitDifference = totalBits - extraBits;
    return 0;
}

/* The allocators HEAD does not call (bitalloc.py:22-125), restated for the widened API surface.
 * mode 0: BitAllocUniform (:22-57, `level` unused); 1: BitAllocConstSNR (:60-91, level[b] = peakSPL for every b);
 * 2: BitAllocConstMNR (:94-125, level = SMR).  The two water-filling loops of the reference do not terminate when bits
 * remain but no band can take another one (every band is full or wider than what is left); that state is detected and
 * reported as -1 (the reference would spin for ever). */
int orc_bitalloc_alt(int mode, double bitBudget, int maxMantBits, int nBands, const int32_t *nLines, const double *level,
                     int32_t *bits)
{
    if (mode == 0) {
        int64_t total = 0;
        for (int b = 0; b < nBands; b++) total += nLines[b];
        int per = (int)(bitBudget / (double)total);                 /* int() truncation, :32 */
        double used = 0;
        for (int b = 0; b < nBands; b++) { bits[b] = per; used += (double)per * nLines[b]; }
        double remaining = bitBudget - used;                        /* :37 */
        if (remaining != 0.0) {
            int64_t line = 0;
            while (remaining > 0) {
                int b = (int)(line % nBands);
                if (nLines[b] == 0 && total == 0) return -1;
                remaining -= nLines[b];
                if (remaining < 0) break;
                if (bits[b] < maxMantBits) bits[b] += 1;
                line++;
            }
        }
    } else {
        double floor_[ORC_MAX_BANDS];
        double remaining = bitBudget;
        for (int b = 0; b < nBands; b++) { bits[b] = 0; floor_[b] = level[b]; }
        while (remaining > 0) {
            int can = 0;
            for (int b = 0; b < nBands; b++) if (bits[b] < maxMantBits && remaining - nLines[b] >= 0) can = 1;
            if (!can) return -1;
            int iMax = 0;
            for (int b = 1; b < nBands; b++) if (floor_[b] > floor_[iMax]) iMax = b;   /* argmax: first index wins */
            if (bits[iMax] < maxMantBits && remaining - nLines[iMax] >= 0) { bits[iMax] += 1; remaining -= nLines[iMax]; }
            floor_[iMax] -= 6.0;
        }
    }
    for (int b = 0; b < nBands; b++) {
        if (bits[b] < 2) bits[b] = 0;                               /* mid-tread: no 1-bit mantissas */
        if (bits[b] > maxMantBits) bits[b] = maxMantBits;
    }
    return 0;
}

/* ------------------------------------------------------------------ masking thresholds */

typedef struct {
    int N, nLines, sampleRate;
    double *zline;   /* Bark(MDCTFreqs)             psychoac.py:95,434 */
    double *tiq;     /* Intensity(Thresh(MDCTFreqs)) psychoac.py:437 */
    double *mld;     /* MLD_F(MDCT_freqs)           psychoac.py:349-372,570-573 */
} PsyTables;

static void psy_tables_init(PsyTables *t, int N, int nLines, int sampleRate)
{
    t->N = N; t->nLines = nLines; t->sampleRate = sampleRate;
    t->zline = (double *)malloc(sizeof(double) * nLines * 3);
    t->tiq = t->zline + nLines;
    t->mld = t->tiq + nLines;
    double mx = 0;
    for (int i = 0; i < nLines; i++) {
        double f = sampleRate / 2.0 / nLines * (i + 0.5);                 /* :434 */
        t->zline[i] = orc_bark(f);
        t->tiq[i] = orc_intensity(orc_thresh(f));
        double f2 = ((i + 0.5) / nLines) * (sampleRate / 2.0);            /* :570 */
        double m = pow(10.0, 1.25 * (1 - cos(M_PI * (fmin(f2, 3000.) / 3000.)) - 2.5));   /* :367 */
        t->mld[i] = m;
        if (m > mx) mx = m;
    }
    for (int i = 0; i < nLines; i++) t->mld[i] /= mx;                     /* :370 */
}

static void psy_tables_free(PsyTables *t) { free(t->zline); }

/* psychoac.py:409-456 (and, with noDrop=0, the identical body of getMaskedThreshold :215-251) */
static void calc_bthr(const PsyTables *t, double *data, int noDrop, double *thr, double *re /*[2N] scratch*/)
{
    int N = t->N, half = N / 2, nLines = t->nLines;
    double *im = re + N;
    orc_hann_window(data, N);                                   /* :428, IN PLACE */
    for (int n = 0; n < N; n++) { re[n] = data[n]; im[n] = 0.0; }
    orc_fft(re, im, N, 0);
    for (int i = 0; i < nLines; i++) thr[i] = 0.0;              /* masked_intensity :431 */
    double *mag = (double *)malloc(sizeof(double) * half);
    for (int k = 0; k < half; k++) mag[k] = hypot(re[k], im[k]);
    int fstep = t->sampleRate / N;                              /* :188 integer division */
    double cnorm = 8.0 / 3.0 * 4.0 / ((double)N * (double)N);   /* :448 */
    double minval = orc_intensity(-30);
    for (int k = 1; k < half - 1; k++) {                        /* findpeaks :166-171 */
        if (!(mag[k] > mag[k - 1] && mag[k] > mag[k + 1] && 10.0 * log10(mag[k]) > -30.0)) continue;
        /* :448  X_fft[index-BW:index+BW] -- python slice semantics: negative start => empty */
        double s = 0.0;
        int lo = k - 3, hi = k + 3;
        if (hi > half) hi = half;
        if (lo >= 0)
            for (int j = lo; j < hi; j++) s += mag[j] * mag[j];
        double inten = cnorm * s;
        if (inten < minval) inten = minval;
        double P = 96 + 10 * log10(inten);
        if (P < -30.) P = -30.;
        /* Masker(f=(k+p)*(sampleRate//N) with p=0, :186-188), :72-85 */
        double zm = orc_bark((double)k * (double)fstep);
        double drop = noDrop ? 0.0 : 15.0;                      /* :83, :450-451 */
        double leveling = 0.367 * (P - 40.0 > 0 ? P - 40.0 : 0.0);   /* :114 */
        for (int i = 0; i < nLines; i++) {                      /* vIntensityAtBark :108-120 */
            double dz = t->zline[i] - zm;
            double a = fabs(dz);
            double spread = ((dz >= 0 ? leveling : 0.0) - 27.0) * ((a - 0.5) * (a > 0.5 ? 1.0 : 0.0));
            double spl = P + spread - drop;
            thr[i] += pow(10.0, (spl - 96) / 10);
        }
    }
    free(mag);
    for (int i = 0; i < nLines; i++) thr[i] = orc_spl(thr[i] + t->tiq[i]);   /* :454-456 */
}

void orc_calc_bthr(double *data, int N, int nMDCTLines, int sampleRate, int noDrop, double *thr)
{
    PsyTables t;
    psy_tables_init(&t, N, nMDCTLines, sampleRate);
    double *scratch = (double *)malloc(sizeof(double) * 2 * N);
    calc_bthr(&t, data, noDrop, thr, scratch);
    free(scratch);
    psy_tables_free(&t);
}

/* psychoac.py:253-318 */
void orc_calc_smrs(double *data, int N, const double *mdct, int nMDCTLines, int mdctScale, int sampleRate,
                   const int32_t *nLines, int nBands, double *smr, double *thr_out)
{
    double *thr = (double *)malloc(sizeof(double) * nMDCTLines);
    orc_calc_bthr(data, N, nMDCTLines, sampleRate, 0, thr);
    int lo = 0;
    double sc = pow(2.0, mdctScale);
    for (int b = 0; b < nBands; b++) {
        smr[b] = 0.0;                                           /* :309 */
        for (int i = lo; i < lo + nLines[b]; i++) {
            double tr = mdct[i] / sc;                           /* :285 */
            double spl = orc_spl(4.0 * (tr * tr));              /* :286-287 */
            double v = spl - thr[i];
            if (i == lo || v > smr[b]) smr[b] = v;
        }
        lo += nLines[b];
    }
    if (thr_out) memcpy(thr_out, thr, sizeof(double) * nMDCTLines);
    free(thr);
}

/* psychoac.py:506-682 */
static void stereo_smr(const PsyTables *t, double *d0, double *d1, const double *X0, const double *X1,
                       const int32_t *scale, const int32_t *nLinesB, int nBands, const int32_t *LRMS,
                       double *smr, double *lines, double *bthr6)
{
    int N = t->N, nL = t->nLines;
    double *buf = (double *)malloc(sizeof(double) * (6 * nL + 2 * N + 2 * N + 4 * nL));
    double *bthr = bthr6 ? bthr6 : buf;            /* L,R,M,S,M',S' */
    double *scratch = buf + 6 * nL;
    double *dm = scratch + 2 * N, *ds = dm + N;
    double *splLR = ds + N;                         /* [2][nL] */
    double *splMS = splLR + 2 * nL;                 /* [2][nL] */
    for (int i = 0; i < nL; i++) {                  /* :534-535 */
        splLR[i] = orc_spl(4. * (X0[i] * X0[i])) - (6.02 * scale[0]);
        splLR[nL + i] = orc_spl(4. * (X1[i] * X1[i])) - (6.02 * scale[1]);
    }
    calc_bthr(t, d0, 0, bthr + 0 * nL, scratch);    /* :540 */
    calc_bthr(t, d1, 0, bthr + 1 * nL, scratch);    /* :541 */
    for (int n = 0; n < N; n++) {                   /* :549 (after the in-place Hann above) */
        dm[n] = (d0[n] + d1[n]) / 2.0;
        ds[n] = (d0[n] - d1[n]) / 2.0;
    }
    double *XM = (double *)malloc(sizeof(double) * 2 * nL), *XS = XM + nL;
    for (int i = 0; i < nL; i++) {                  /* :551 */
        XM[i] = (X0[i] + X1[i]) / 2.0;
        XS[i] = (X0[i] - X1[i]) / 2.0;
        splMS[i] = orc_spl(4. * (XM[i] * XM[i])) - (6.02 * scale[0]);        /* :554 */
        splMS[nL + i] = orc_spl(4. * (XS[i] * XS[i])) - (6.02 * scale[1]);   /* :555 */
    }
    calc_bthr(t, dm, 0, bthr + 2 * nL, scratch);    /* :559 */
    calc_bthr(t, ds, 0, bthr + 3 * nL, scratch);    /* :560 */
    calc_bthr(t, dm, 1, bthr + 4 * nL, scratch);    /* :561 (third Hann on the same array) */
    calc_bthr(t, ds, 1, bthr + 5 * nL, scratch);    /* :562 */
    int lo = 0;
    for (int b = 0; b < nBands; b++) {
        double sLR[2] = {-96.0, -96.0}, sMS[2] = {-96.0, -96.0};   /* :496-498 */
        for (int i = lo; i < lo + nLinesB[b]; i++) {
            double M = bthr[2 * nL + i], S = bthr[3 * nL + i];
            double mldM = bthr[4 * nL + i] * t->mld[i];           /* :582 */
            double mldS = bthr[5 * nL + i] * t->mld[i];           /* :583 */
            double thrM = fmax(M, fmin(S, mldS));                 /* :591 */
            double thrS = fmax(S, fmin(M, mldM));
            double v;
            v = splLR[i] - bthr[i];           if (i == lo || v > sLR[0]) sLR[0] = v;   /* :594 */
            v = splLR[nL + i] - bthr[nL + i]; if (i == lo || v > sLR[1]) sLR[1] = v;
            v = splMS[i] - thrM;              if (i == lo || v > sMS[0]) sMS[0] = v;   /* :597 */
            v = splMS[nL + i] - thrS;         if (i == lo || v > sMS[1]) sMS[1] = v;
        }
        for (int ch = 0; ch < 2; ch++) {                          /* :665-680 */
            smr[ch * nBands + b] = LRMS[b] ? sMS[ch] : sLR[ch];
            for (int i = lo; i < lo + nLinesB[b]; i++)
                lines[ch * nL + i] = LRMS[b] ? (ch ? XS[i] : XM[i]) : (ch ? X1[i] : X0[i]);
        }
        lo += nLinesB[b];
    }
    free(XM);
    free(buf);
}

void orc_stereo_smr(double *data0, double *data1, int N, const double *mdct0, const double *mdct1,
                    int nMDCTLines, const int32_t *scale, int sampleRate, const int32_t *nLines, int nBands,
                    const int32_t *LRMS, double *smr, double *lines, double *bthr6)
{
    PsyTables t;
    psy_tables_init(&t, N, nMDCTLines, sampleRate);
    stereo_smr(&t, data0, data1, mdct0, mdct1, scale, nLines, nBands, LRMS, smr, lines, bthr6);
    psy_tables_free(&t);
}

/* codec.py:96-102 */
void orc_lrms(const double *l, const double *r, int N, const int32_t *nLines, int nBands, int32_t *LRMS)
{
    double *buf = (double *)malloc(sizeof(double) * 4 * N);
    double *lr = buf, *li = buf + N, *rr = buf + 2 * N, *ri = buf + 3 * N;
    for (int n = 0; n < N; n++) { lr[n] = l[n]; li[n] = 0; rr[n] = r[n]; ri[n] = 0; }
    orc_fft(lr, li, N, 0);
    orc_fft(rr, ri, N, 0);
    int lo = 0;
    for (int b = 0; b < nBands; b++) {
        double dr = 0, di = 0, sr = 0, si = 0;
        for (int k = lo; k < lo + nLines[b]; k++) {
            double l2r = lr[k] * lr[k] - li[k] * li[k], l2i = lr[k] * li[k] + li[k] * lr[k];
            double r2r = rr[k] * rr[k] - ri[k] * ri[k], r2i = rr[k] * ri[k] + ri[k] * rr[k];
            dr += l2r - r2r; di += l2i - r2i;
            sr += l2r + r2r; si += l2i + r2i;
        }
        LRMS[b] = hypot(dr, di) < 0.8 * hypot(sr, si);
        lo += nLines[b];
    }
    free(buf);
}

/* ------------------------------------------------------------------ bit writer / reader */

typedef struct { uint8_t *data; int64_t cap; int64_t bitpos; } BitW;

/* bitpack.py:36-101: MSB-first, lowest nBits of info */
static void bw_write(BitW *w, uint64_t info, int nBits)
{
    for (int i = nBits - 1; i >= 0; i--) {
        int64_t byte = w->bitpos >> 3;
        if (byte < w->cap && ((info >> i) & 1)) w->data[byte] |= (uint8_t)(0x80 >> (w->bitpos & 7));
        w->bitpos++;
    }
}

typedef struct { const uint8_t *data; int64_t nbytes; int64_t bitpos; int err; } BitR;

/* bitpack.py:104-170 */
static uint32_t br_read(BitR *r, int nBits)
{
    uint32_t v = 0;
    for (int i = 0; i < nBits; i++) {
        int64_t byte = r->bitpos >> 3;
        int bit = 0;
        if (byte < r->nbytes) bit = (r->data[byte] >> (7 - (r->bitpos & 7))) & 1;
        else r->err = 1;
        v = (v << 1) | (uint32_t)bit;
        r->bitpos++;
    }
    return v;
}

/* ------------------------------------------------------------------ Huffman */

static inline int huff_len(const OrcHuff *h, int t, uint32_t v, int ba)
{
    if ((int64_t)v < h->nkeys[t]) {
        int l = h->len[h->off[t] + v];
        if (l) return l;
    }
    return h->esc_len[t] + ba;        /* Huffman.py:296-298 */
}

/* ------------------------------------------------------------------ whole-stream encode */

int64_t orc_encoded_blocks(int64_t nSamples, int nMDCTLines)
{
    /* pcmfile.py:66-82 (ceil(n/1024) data blocks) + pacfile.py:355-365 (one flush block) */
    return (nSamples + nMDCTLines - 1) / nMDCTLines + 1;
}

static void put_le32(uint8_t *p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }
static void put_le16(uint8_t *p, uint32_t v) { p[0] = v; p[1] = v >> 8; }
static uint32_t get_le32(const uint8_t *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint32_t get_le16(const uint8_t *p) { return p[0] | (p[1] << 8); }

int64_t orc_encode_stream(const OrcParams *p, const OrcHuff *h, const int16_t *pcm, int64_t nSamples,
                          uint8_t *out, int64_t cap, OrcTrace *tr, int64_t *final_state)
{
    if (p->nChannels != 2 || p->nMDCTLines <= 0 || (p->nMDCTLines & (p->nMDCTLines - 1))) return -1;
    const int halfN = p->nMDCTLines, N = 2 * halfN;
    int32_t nLines[ORC_MAX_BANDS], lower[ORC_MAX_BANDS];
    const int nBands = orc_band_layout(halfN, p->sampleRate, nLines);
    for (int b = 0, lo = 0; b < nBands; b++) { lower[b] = lo; lo += nLines[b]; }

    /* ---- header, pacfile.py:231-261 ---- */
    int64_t pos = 0;
    int64_t hdr = 4 + 18 + 4 + 2 * nBands;
    if (cap < hdr) return -2;
    memcpy(out, "PAC ", 4);
    uint32_t numSamplesHdr = (uint32_t)nSamples;
    if (nSamples % halfN == 0) numSamplesHdr += halfN;          /* :240-242 (inverted padding rule) */
    put_le32(out + 4, (uint32_t)p->sampleRate);
    put_le16(out + 8, (uint32_t)p->nChannels);
    put_le32(out + 10, numSamplesHdr);
    put_le32(out + 14, (uint32_t)halfN);
    put_le16(out + 18, (uint32_t)p->nScaleBits);
    put_le16(out + 20, (uint32_t)p->nMantSizeBits);
    put_le32(out + 22, (uint32_t)nBands);
    for (int b = 0; b < nBands; b++) put_le16(out + 26 + 2 * b, (uint32_t)nLines[b]);
    pos = hdr;

    PsyTables T;
    psy_tables_init(&T, N, halfN, p->sampleRate);
    double *full = (double *)malloc(sizeof(double) * (2 * N + 2 * N + 2 * halfN + 2 * halfN + 2 * nBands));
    double *raw = full + 2 * N;               /* un-windowed copy for the LRMS FFTs */
    double *mdct = raw + 2 * N;               /* [2][halfN] */
    double *lines = mdct + 2 * halfN;         /* [2][halfN] */
    double *smr = lines + 2 * halfN;          /* [2][nBands] */
    uint64_t *mant = (uint64_t *)malloc(sizeof(uint64_t) * 2 * halfN);   /* signed codes per line position */
    int64_t extraBits = 0, bitDeposit = 0;    /* pacfile.py:269, Huffman.py:262 */
    const int maxMantBits = (1 << p->nMantSizeBits) > 16 ? 16 : (1 << p->nMantSizeBits);   /* codec.py:218-219 */
    const int64_t nBlocks = orc_encoded_blocks(nSamples, halfN);
    int64_t status = 0;

    for (int64_t blk = 0; blk < nBlocks; blk++) {
        /* ---- PCM -> signed fractions, pcmfile.py:66-100; prior|current, pacfile.py:279-282 ---- */
        for (int ch = 0; ch < 2; ch++)
            for (int n = 0; n < N; n++) {
                int64_t s = (blk - 1) * (int64_t)halfN + n;
                double v = 0.0;
                if (s >= 0 && s < nSamples) {
                    int c = pcm[2 * s + ch];
                    int neg = c < 0;
                    int64_t code = neg ? -(int64_t)c : c;       /* pcmfile.py:92-93 */
                    if (code & 32768) code -= 32768;            /* quantize.py:133-138 (-32768 -> 0) */
                    v = 2.0 * (double)code / 65535.0;           /* quantize.py:141 */
                    if (neg) v = -v;
                }
                full[ch * N + n] = v;
                raw[ch * N + n] = v;
            }
        /* ---- codec.Encode, codec.py:83-129 ---- */
        int32_t LRMS[ORC_MAX_BANDS];
        orc_lrms(raw, raw + N, N, nLines, nBands, LRMS);        /* :96-102 */

        /* ---- EncodeDualChannel, codec.py:212-281 ---- */
        double bitBudget = p->targetBitsPerSample * halfN;      /* :223 */
        bitBudget -= p->nScaleBits * (nBands + 1);              /* :224 */
        bitBudget -= p->nMantSizeBits * nBands;                 /* :225 */
        bitBudget -= p->nTableIDBits;                           /* :227 */
        {   /* huffman.withdrawBits(), Huffman.py:363-371 */
            int64_t extra = 0;
            if (bitDeposit > 10) {
                extra = bitDeposit / 100;       /* floor == trunc for positives */
                bitDeposit -= extra;
            } else if (bitDeposit < 0) {
                extra = bitDeposit;
                bitDeposit = 0;
            }
            extraBits += extra;                                 /* codec.py:229 */
        }
        int32_t overallScale[2];
        for (int ch = 0; ch < 2; ch++) {                        /* :237-246 */
            if (p->window == 1) {                               /* KBDWindow returns a copy (window.py:64): `full` stays un-windowed */
                double *tmpw = (double *)malloc((size_t)N * sizeof(double));
                orc_kbd_window(full + ch * N, tmpw, N, 4.0);
                orc_mdct(tmpw, halfN, halfN, mdct + ch * halfN);
                free(tmpw);
            } else {
                orc_sine_window(full + ch * N, N);              /* in place: the psychoacoustic model sees it (window.py:37) */
                orc_mdct(full + ch * N, halfN, halfN, mdct + ch * halfN);
            }
            double maxLine = 0.0;
            for (int i = 0; i < halfN; i++) if (fabs(mdct[ch * halfN + i]) > maxLine) maxLine = fabs(mdct[ch * halfN + i]);
            overallScale[ch] = orc_scale_factor(maxLine, p->nScaleBits, 5);
            double g = (double)(1 << overallScale[ch]);
            for (int i = 0; i < halfN; i++) mdct[ch * halfN + i] *= g;
        }
        stereo_smr(&T, full, full + N, mdct, mdct + halfN, overallScale, nLines, nBands, LRMS, smr, lines, NULL);  /* :250 */

        int32_t ba[2][ORC_MAX_BANDS], sf[2][ORC_MAX_BANDS];
        for (int ch = 0; ch < 2; ch++) {                        /* :257-277 */
            int64_t diff;
            orc_bitalloc(bitBudget, extraBits, maxMantBits, nBands, nLines, smr + ch * nBands, LRMS, ba[ch], &diff);
            extraBits += diff;
            for (int b = 0; b < nBands; b++) {
                double scaleLine = 0.0;
                const double *x = lines + ch * halfN + lower[b];
                for (int i = 0; i < nLines[b]; i++) if (fabs(x[i]) > scaleLine) scaleLine = fabs(x[i]);
                sf[ch][b] = orc_scale_factor(scaleLine, p->nScaleBits, ba[ch][b]);
                if (ba[ch][b]) orc_vmantissa(x, nLines[b], sf[ch][b], p->nScaleBits, ba[ch][b], mant + ch * halfN + lower[b]);
                else for (int i = 0; i < nLines[b]; i++) mant[ch * halfN + lower[b] + i] = 0;
            }
        }
        /* ---- sign strip + Huffman table search + deposit, codec.py:111-125, Huffman.py:274-309 ---- */
        int tableID[2];
        int64_t huffBits[2];
        for (int ch = 0; ch < 2; ch++) {
            int64_t best = 0;
            int bestID = 1;
            int64_t nMant = 0, origin = 0;
            for (int b = 0; b < nBands; b++) if (ba[ch][b]) { nMant += nLines[b]; origin += (int64_t)ba[ch][b] * nLines[b]; }
            for (int t = 0; t < ORC_NTABLES; t++) {
                int64_t tot = 0;
                for (int b = 0; b < nBands; b++) {
                    if (!ba[ch][b]) continue;
                    uint64_t mask = ((uint64_t)1 << (ba[ch][b] - 1)) - 1;
                    for (int i = 0; i < nLines[b]; i++)
                        tot += huff_len(h, t, (uint32_t)(mant[ch * halfN + lower[b] + i] & mask), ba[ch][b]);
                }
                if (t == 0 || tot < best) { best = tot; bestID = t + 1; }   /* Huffman.py:300-307 */
            }
            tableID[ch] = bestID;
            huffBits[ch] = best;
            bitDeposit += origin - (best + nMant + p->nTableIDBits);       /* codec.py:118-120 */
        }
        /* ---- WriteDataBlock, pacfile.py:288-351 ---- */
        for (int ch = 0; ch < 2; ch++) {
            int64_t nbits = p->nScaleBits + p->nTableIDBits;
            for (int b = 0; b < nBands; b++) {
                nbits += p->nMantSizeBits + p->nScaleBits;
                if (ba[ch][b]) nbits += nLines[b];
            }
            nbits += huffBits[ch];
            nbits += nBands;                                     /* :312 LRMS */
            int64_t nBytes = (nbits + 7) / 8;                    /* :315-316 */
            if (pos + 4 + nBytes > cap) { status = -2; goto done; }
            put_le32(out + pos, (uint32_t)nBytes);
            pos += 4;
            memset(out + pos, 0, (size_t)nBytes);
            BitW w = {out + pos, nBytes, 0};
            bw_write(&w, (uint64_t)overallScale[ch], p->nScaleBits);
            bw_write(&w, (uint64_t)tableID[ch], p->nTableIDBits);
            int t = tableID[ch] - 1;
            for (int b = 0; b < nBands; b++) {
                int a = ba[ch][b];
                bw_write(&w, (uint64_t)(a ? a - 1 : 0), p->nMantSizeBits);   /* :329-331 */
                bw_write(&w, (uint64_t)sf[ch][b], p->nScaleBits);           /* :332 */
                if (!a) continue;
                const uint64_t *m = mant + ch * halfN + lower[b];
                uint64_t mask = ((uint64_t)1 << (a - 1)) - 1;
                for (int i = 0; i < nLines[b]; i++) bw_write(&w, (m[i] >> (a - 1)) & 1, 1);   /* :335-336 */
                for (int i = 0; i < nLines[b]; i++) {                                        /* :337-341 */
                    uint32_t v = (uint32_t)(m[i] & mask);
                    int l = (int64_t)v < h->nkeys[t] ? h->len[h->off[t] + v] : 0;
                    if (l) bw_write(&w, h->code[h->off[t] + v], l);
                    else { bw_write(&w, h->esc_code[t], h->esc_len[t]); bw_write(&w, v, a); }
                }
            }
            for (int b = 0; b < nBands; b++) bw_write(&w, (uint64_t)LRMS[b], 1);             /* :347-348 */
            pos += nBytes;
            if (tr && tr->nbytes) tr->nbytes[blk * 2 + ch] = (int32_t)nBytes;
        }
        if (tr) {
            if (tr->lrms) { int32_t m = 0; for (int b = 0; b < nBands; b++) m |= (LRMS[b] ? 1 : 0) << b; tr->lrms[blk] = m; }
            for (int ch = 0; ch < 2; ch++) {
                if (tr->oscale) tr->oscale[blk * 2 + ch] = overallScale[ch];
                if (tr->tableID) tr->tableID[blk * 2 + ch] = tableID[ch];
                for (int b = 0; b < nBands; b++) {
                    if (tr->ba) tr->ba[(blk * 2 + ch) * nBands + b] = ba[ch][b];
                    if (tr->sf) tr->sf[(blk * 2 + ch) * nBands + b] = sf[ch][b];
                    if (tr->smr) tr->smr[(blk * 2 + ch) * nBands + b] = smr[ch * nBands + b];
                }
                if (tr->lines) memcpy(tr->lines + (blk * 2 + ch) * halfN, lines + ch * halfN, sizeof(double) * halfN);
                if (tr->mant) for (int i = 0; i < halfN; i++) tr->mant[(blk * 2 + ch) * halfN + i] = (int32_t)mant[ch * halfN + i];
            }
            if (tr->extraBits) tr->extraBits[blk] = extraBits;
            if (tr->bitDeposit) tr->bitDeposit[blk] = bitDeposit;
        }
    }
done:
    if (final_state) { final_state[0] = bitDeposit; final_state[1] = extraBits; }
    free(mant);
    free(full);
    psy_tables_free(&T);
    return status < 0 ? status : pos;
}

/* ------------------------------------------------------------------ whole-stream decode */

/* Huffman.py:321-344: bit-serial prefix match against decodingTable.  Built as a binary trie. */
typedef struct { int32_t child[2]; int32_t sym; } TrieNode;   /* sym: -2 none, -1 escape, >=0 magnitude */
typedef struct { TrieNode *n; int count, cap; } Trie;

static int trie_new(Trie *t)
{
    if (t->count == t->cap) { t->cap = t->cap ? t->cap * 2 : 1024; t->n = (TrieNode *)realloc(t->n, sizeof(TrieNode) * t->cap); }
    t->n[t->count].child[0] = t->n[t->count].child[1] = -1;
    t->n[t->count].sym = -2;
    return t->count++;
}

static void trie_add(Trie *t, uint32_t code, int len, int sym)
{
    int cur = 0;
    for (int i = len - 1; i >= 0; i--) {
        int bit = (code >> i) & 1;
        if (t->n[cur].child[bit] < 0) { int c = trie_new(t); t->n[cur].child[bit] = c; }
        cur = t->n[cur].child[bit];
    }
    t->n[cur].sym = sym;
}

static void trie_build(Trie *t, const OrcHuff *h, int tab)
{
    t->n = NULL; t->count = t->cap = 0;
    trie_new(t);
    for (int v = 0; v < h->nkeys[tab]; v++)
        if (h->len[h->off[tab] + v]) trie_add(t, h->code[h->off[tab] + v], h->len[h->off[tab] + v], v);
    trie_add(t, h->esc_code[tab], h->esc_len[tab], -1);
}

int64_t orc_decode_stream(const OrcHuff *h, const uint8_t *pac, int64_t nbytes, int16_t *pcm, int64_t capSamples,
                          OrcParams *hdr_out, int64_t *numSamplesHdr)
{
    return orc_decode_stream_w(h, pac, nbytes, pcm, capSamples, hdr_out, numSamplesHdr, 0);
}

int64_t orc_decode_stream_w(const OrcHuff *h, const uint8_t *pac, int64_t nbytes, int16_t *pcm, int64_t capSamples,
                            OrcParams *hdr_out, int64_t *numSamplesHdr, int window)
{
    /* ---- header, pacfile.py:123-151 ---- */
    if (nbytes < 26 || memcmp(pac, "PAC ", 4)) return -1;
    OrcParams p;
    p.sampleRate = (int32_t)get_le32(pac + 4);
    p.nChannels = (int32_t)get_le16(pac + 8);
    uint32_t numSamples = get_le32(pac + 10);
    p.nMDCTLines = (int32_t)get_le32(pac + 14);
    p.nScaleBits = (int32_t)get_le16(pac + 18);
    p.nMantSizeBits = (int32_t)get_le16(pac + 20);
    p.nTableIDBits = 4;                                          /* pacfile.py:189 */
    p.window = window; p.reserved = 0;
    p.targetBitsPerSample = 0;
    int nBands = (int)get_le32(pac + 22);
    if (p.nChannels != 2 || nBands > ORC_MAX_BANDS || nbytes < 26 + 2 * nBands) return -1;
    int32_t nLines[ORC_MAX_BANDS], lower[ORC_MAX_BANDS];
    for (int b = 0, lo = 0; b < nBands; b++) { nLines[b] = (int32_t)get_le16(pac + 26 + 2 * b); lower[b] = lo; lo += nLines[b]; }
    if (hdr_out) *hdr_out = p;
    if (numSamplesHdr) *numSamplesHdr = numSamples;
    const int halfN = p.nMDCTLines, N = 2 * halfN;
    int64_t pos = 26 + 2 * nBands;

    Trie tries[ORC_NTABLES];
    for (int t = 0; t < ORC_NTABLES; t++) trie_build(&tries[t], h, t);

    double *ola = (double *)calloc((size_t)(2 * halfN + 2 * halfN + 2 * N), sizeof(double));
    double *line = ola + 2 * halfN;          /* [2][halfN] */
    double *dec = line + 2 * halfN;          /* [2][N] */
    int64_t *mant = (int64_t *)malloc(sizeof(int64_t) * 2 * halfN);
    int64_t nOut = 0, status = 0;
    int first = 1;

    for (;;) {
        int32_t ba[2][ORC_MAX_BANDS], sf[2][ORC_MAX_BANDS], oscale[2], LRMS[ORC_MAX_BANDS];
        int eof = 0;
        for (int i = 0; i < 2 * halfN; i++) mant[i] = 0;
        for (int ch = 0; ch < 2; ch++) {                         /* pacfile.py:167-217 */
            if (pos + 4 > nbytes) { eof = 1; break; }            /* :170-178 */
            int64_t nB = get_le32(pac + pos);
            pos += 4;
            if (pos + nB > nbytes) { status = -3; goto done; }   /* :184 */
            BitR r = {pac + pos, nB, 0, 0};
            pos += nB;
            oscale[ch] = (int32_t)br_read(&r, p.nScaleBits);
            int tid = (int)br_read(&r, p.nTableIDBits);
            if (tid < 1 || tid > ORC_NTABLES) { status = -4; goto done; }
            const Trie *T = &tries[tid - 1];
            for (int b = 0; b < nBands; b++) {
                int a = (int)br_read(&r, p.nMantSizeBits);
                if (a) a += 1;                                   /* :196 */
                ba[ch][b] = a;
                sf[ch][b] = (int32_t)br_read(&r, p.nScaleBits);
                if (!a) continue;
                int64_t *m = mant + ch * halfN + lower[b];
                for (int i = 0; i < nLines[b]; i++) m[i] = (int64_t)br_read(&r, 1) << (a - 1);   /* :203-204,210 */
                for (int i = 0; i < nLines[b]; i++) {
                    int cur = 0;
                    while (T->n[cur].sym == -2) {                /* Huffman.py:337-344 */
                        int bit = (int)br_read(&r, 1);
                        cur = T->n[cur].child[bit];
                        if (cur < 0 || r.err) { status = -5; goto done; }
                    }
                    int64_t v = T->n[cur].sym;
                    if (v == -1) v = br_read(&r, a);             /* Huffman.py:326-327 */
                    m[i] += v;
                }
            }
            for (int b = 0; b < nBands; b++) LRMS[b] = (int32_t)br_read(&r, 1);   /* :216-217 */
        }
        if (eof) break;
        /* ---- codec.Decode, codec.py:25-65 ---- */
        for (int ch = 0; ch < 2; ch++) {
            double rescale = 1. * (double)(1 << oscale[ch]);
            for (int i = 0; i < halfN; i++) line[ch * halfN + i] = 0.0;
            for (int b = 0; b < nBands; b++)
                if (ba[ch][b]) orc_vdequantize(sf[ch][b], mant + ch * halfN + lower[b], nLines[b], p.nScaleBits, ba[ch][b], line + ch * halfN + lower[b]);
            for (int i = 0; i < halfN; i++) line[ch * halfN + i] /= rescale;   /* :43 */
        }
        for (int b = 0; b < nBands; b++) {
            if (!LRMS[b]) continue;
            for (int i = lower[b]; i < lower[b] + nLines[b]; i++) {
                /* :46-56: mdctLineL aliases mdctLine[0], so R is built from the already-updated L */
                line[i] = line[i] - line[halfN + i];
                line[halfN + i] = line[i] + line[halfN + i];
            }
        }
        for (int ch = 0; ch < 2; ch++) {                         /* :59-60 */
            orc_imdct(line + ch * halfN, halfN, halfN, dec + ch * N);
            if (window == 1) {
                double *tmpw = (double *)malloc((size_t)N * sizeof(double));
                orc_kbd_window(dec + ch * N, tmpw, N, 4.0);
                memcpy(dec + ch * N, tmpw, (size_t)N * sizeof(double));
                free(tmpw);
            } else orc_sine_window(dec + ch * N, N);
        }
        /* overlap-add, pacfile.py:223-226; __main__ drops the first block, pacfile.py:485-487 */
        if (!first) {
            if (nOut + halfN > capSamples) { status = -2; goto done; }
            for (int i = 0; i < halfN; i++)
                for (int ch = 0; ch < 2; ch++) {
                    double v = ola[ch * halfN + i] + dec[ch * N + i];
                    /* pcmfile.py:127-134 + quantize.py:91-117 with 16 bits */
                    double a = fabs(v);
                    int64_t code = a < 1 ? (int64_t)((a * 65535.0 + 1.0) / 2.0) : 32767;
                    pcm[2 * (nOut + i) + ch] = (int16_t)(signbit(v) ? -code : code);
                }
            nOut += halfN;
        }
        first = 0;
        for (int ch = 0; ch < 2; ch++) memcpy(ola + ch * halfN, dec + ch * N + halfN, sizeof(double) * halfN);
    }
    /* EOF: the saved overlap tail is returned once, pacfile.py:171-176 */
    if (nOut + halfN > capSamples) { status = -2; goto done; }
    for (int i = 0; i < halfN; i++)
        for (int ch = 0; ch < 2; ch++) {
            double v = ola[ch * halfN + i];
            double a = fabs(v);
            int64_t code = a < 1 ? (int64_t)((a * 65535.0 + 1.0) / 2.0) : 32767;
            pcm[2 * (nOut + i) + ch] = (int16_t)(signbit(v) ? -code : code);
        }
    nOut += halfN;
done:
    for (int t = 0; t < ORC_NTABLES; t++) free(tries[t].n);
    free(mant);
    free(ola);
    return status < 0 ? status : nOut;
}

/* ------------------------------------------------------------------ threaded batch (CPU baseline) */

int orc_encode_batch(const OrcParams *p, const OrcHuff *h, const int16_t *pcm, int64_t nSamples, int S,
                     uint8_t *out, int64_t cap, int64_t *outBytes, int nthreads)
{
    int rc = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int s = 0; s < S; s++) {
        int64_t n = orc_encode_stream(p, h, pcm + (int64_t)s * nSamples * 2, nSamples, out + (int64_t)s * cap, cap, NULL, NULL);
        outBytes[s] = n;
        if (n < 0) {
#pragma omp critical
            rc = (int)n;
        }
    }
    return rc;
}
