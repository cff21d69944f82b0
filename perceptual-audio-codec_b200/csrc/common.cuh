// common.cuh -- shared types and device helpers for the PAC B200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pac_b200.h"

namespace pac {

constexpr int kMaxBands = PAC_MAX_BANDS;
constexpr int kNTables = PAC_NTABLES;
constexpr int kLenLutSize = 17920;   // >= max Huffman key + 1 (17906), padded

template <typename T> struct Vec2;
template <> struct Vec2<float>  { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <typename T> __host__ __device__ inline typename Vec2<T>::type mk2(T a, T b) {
    typename Vec2<T>::type r; r.x = a; r.y = b; return r;
}
template <typename V> __device__ inline V cmul(V a, V b) { V r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <typename V> __device__ inline V cadd(V a, V b) { V r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename V> __device__ inline V csub(V a, V b) { V r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename V> __device__ inline V cconj(V a) { a.y = -a.y; return a; }

// Static band layout (psychoac.py:124-156,193-213), passed by value to kernels.
struct BandInfo {
    int32_t nBands;
    int16_t lo[kMaxBands + 1];   // lo[b] .. lo[b+1]-1 are the lines of band b
};

// Per-(sampleRate,N) constant tables in global memory (all computed on the host in double, libm).
template <typename T>
struct DevTables {
    using T2 = typename Vec2<T>::type;
    const T *sinw;        // [N]   the MDCT window: sin((n+.5)pi/N) window.py:35-37, or the KBD window (window.py:56-78) with PAC_WINDOW_KBD
    const T *hann;        // [N]   .5(1-cos(2pi(n+.5)/N))                window.py:49-51
    const T2 *tw;         // [M]   exp(-2pi i m/M), M = N/2: twiddles of the M- and M/2-point FFTs
    const T2 *tw_split;   // [M+1] exp(-2pi i k/N): real-FFT split
    const T2 *mdct_pre;   // [M/2] exp(-i pi n/M)
    const T2 *mdct_post;  // [M/2] exp(-i pi (n+1/4)/M)
    const T *zline;       // [M]   Bark((i+.5) fs/2/M)                   psychoac.py:95,434
    const T *tiq;         // [M]   Intensity(Thresh(f_i))                psychoac.py:437
    const T *mld;         // [M]   MLD_F                                 psychoac.py:349-372,570-573
    const T *zpeak;       // [M]   Bark(k * (fs // N))                   psychoac.py:186-188 (Python-2 int division)
    const uint8_t *band_of_line;   // [M]
    T2 hann_w;            // exp(i pi/N): frequency-domain Hann taps
    T cnorm;              // 8/3*4/N^2                                   psychoac.py:448
    T imdct_scale;
    int winInPlace;       // 1: the window multiplies the block in place, so the psychoacoustic model sees windowed data (SineWindow,
                          //    SURVEY App. A Q1); 0: it is applied to a copy (KBDWindow) and only the MDCT sees it
};

// Static geometry + scan weights of the fp32 fast threshold evaluation (analysis.cuh: masked_curve_fast; numpy model
// in tests/model_analysis.py: Geometry / weighted_suffix_scan / curve_v3).  The fixed-slope skirts are scans over the BINS
// (Bark positions zp_k of the masker frequencies k * (fs // N)): NT = M/4 threads own 4 consecutive bins each, a warp 128.
struct FastTables {
    const uint2 *binEU;          // [NT] per thread: first line of the upper skirt (M: none) of its four bins, 16 bits each
    const float2 *binZ;          // [M] per bin: Bark position of the masker frequency as hi + lo floats
    const float4 *lineRec;       // [M] per line: (pa | pb << 16 as bits, fL, fU, threshold in quiet) -- plateau bins [pa, pb);
                                 //     lower skirts = SD[min(pb, M-1)] * fL, quiet upper skirts = SA[max(pa-1, 0)] * fU
    const float2 *lineZ;         // [M] per line: Bark position as hi + lo floats
    const short *kcountU;        // [M+1] kcountU[i] = number of bins whose upper skirt starts at or before line i
    const float *sD;             // [13][NT] descending scan weights: wl[3], ww[5], wc, wf[4]
    const float *sA;             // [10][NT] ascending scan weights: ww[5], wc, wf[4]
    float omD[16], omA[16];      // per-warp carry weights
    uint32_t kUhc[32];           // per 64-line half-chunk h: kcountU[64h] | kcountU[64h + 63] << 16 (read from the constant bank where it is needed:
                                 // as per-thread registers carried through the curves these were spilled and re-read from local memory)
};

// Encoder scalars derived from PacParams
struct EncConsts {
    double bitBudget;      // codec.py:223-227
    int32_t nScaleBits, nMantSizeBits, nTableIDBits, maxMantBits;
    int32_t fixedBits;     // per-channel chunk bits that do not depend on the data (pacfile.py:291-296,312)
    int32_t esc_len[kNTables];
    uint32_t esc_code[kNTables];
    int32_t nkeys[kNTables];
    int32_t off[kNTables];
};

// ------------------------------------------------------------------ quantiser (quantize.py), always in double
// QuantizeUniform of |x| with R bits, without the sign (quantize.py:40-64 / :91-117)
__device__ __forceinline__ unsigned long long quant_mag(double a, int R) {
    unsigned long long half = 1ull << (R - 1);
    if (a >= 1.0) return half - 1ull;
    double largest = (double)(half << 1) - 1.0;
    return (unsigned long long)((a * largest + 1.0) / 2.0);
}

// ScaleFactor(aNum, nScaleBits, nMantBits) (quantize.py:148-177)
__device__ __forceinline__ int scale_factor(double a, int nScaleBits, int nMantBits) {
    if (nMantBits <= 0) return 0;
    if (nScaleBits < 0) nScaleBits = 0;
    int largestScale = (1 << nScaleBits) - 1;
    int R = nMantBits + largestScale;
    unsigned long long q = quant_mag(fabs(a), R) << 1;
    if (q == 0) return largestScale;
    int p = 63 - __clzll((long long)q);       // msb position of (q<<1)
    int z = (R - 1) - p;                       // shifts until bit R-1 is set
    if (z < 0) z = 0;
    return z < largestScale ? z : largestScale;
}

// magnitude part of vMantissa (quantize.py:315-342): (q << (scale+1)) >> (R - nMantBits + 1)
__device__ __forceinline__ unsigned int mant_mag(double a, int scale, int largestScale, int nMantBits) {
    int R = nMantBits + largestScale;
    unsigned long long q = quant_mag(a, R);
    return (unsigned int)((q << (scale + 1)) >> (R - nMantBits + 1));
}

// vDequantize of one code (quantize.py:345-376)
__device__ __forceinline__ double dequant(int scale, long long m, int largestScale, int nMantBits) {
    long long signMask = 1ll << (nMantBits - 1);
    int R = nMantBits + largestScale;
    bool neg = (m & signMask) == signMask;
    if (neg) m -= signMask;
    long long q = m << (largestScale - scale);
    if (scale < largestScale && m > 0) q += 1ll << (largestScale - scale - 1);
    double largest = (double)(1ll << R) - 1.0;
    double a = 2.0 * (double)q / largest;
    return neg ? -a : a;
}

// ------------------------------------------------------------------ warp helpers
__device__ __forceinline__ unsigned long long sortable(double v) {
    unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { T w = __shfl_xor_sync(0xffffffffu, v, o); v = w > v ? w : v; }
    return v;
}
template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------ math per precision
template <typename T> struct M;
template <> struct M<double> {
    __device__ static __forceinline__ double log10_(double x) { return log10(x); }
    // Intensity(spl) = 10**((spl-96)/10)   psychoac.py:37-42
    __device__ static __forceinline__ double intensity(double spl) { return exp10((spl - 96.0) / 10.0); }
    __device__ static __forceinline__ double min_intensity() { return 0x1.1ad032752ca29p-42; }   // 10**-12.6 as Python computes it (psychoac.py:20)
};
template <> struct M<float> {
    __device__ static __forceinline__ float log10_(float x) { return log10f(x); }
    __device__ static __forceinline__ float intensity(float spl) { return exp2f((spl - 96.0f) * 0.33219280948873623f); }
    __device__ static __forceinline__ float min_intensity() { return 2.511886431509582e-13f; }
};

// SPL(intensity)  psychoac.py:15-35
template <typename T> __device__ __forceinline__ T spl_of(T inten) {
    T mn = M<T>::min_intensity();
    if (inten < mn) inten = mn;
    T s = (T)96 + (T)10 * M<T>::log10_(inten);
    return s < (T)-30 ? (T)-30 : s;
}

}  // namespace pac
