// stages.cuh -- kernels behind the reference's L2 function names (window.py, mdct.py, quantize.py, bitalloc.py)
// when they are called one stage at a time through the Python shim.
#pragma once
#include "common.cuh"
#include "fft.cuh"
#include "scan.cuh"

namespace pac {

// window.py:27-53,56-78.  w[n] precomputed on the host (sine / hann / kbd), x [n][N] in place.
__global__ void k_window(double *x, const double *w, int64_t total, int N) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        x[i] *= w[i % N];
}

template <typename T, int LOGM>
struct MdctSmem {
    static constexpr int M = 1 << LOGM;
    using T2 = typename Vec2<T>::type;
    T x[2 * M];
    T2 W[M / 2 + 2];
};

// mdct.py:49-71: MDCT(data, M, M) by fold + M/2-point complex FFT (tests/model_analysis.py:mdct_fold_fft)
template <typename T, int LOGM>
__global__ void k_mdct(const double *in, double *out, int n, DevTables<T> tb) {
    using SS = MdctSmem<T, LOGM>;
    using T2 = typename Vec2<T>::type;
    constexpr int M = SS::M, N = 2 * M, H = M / 2, NT = (M / 4 > 0 ? M / 4 : 1);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SS &sm = *reinterpret_cast<SS *>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t w = blockIdx.x;
    for (int i = tid; i < N; i += NT) sm.x[i] = (T)in[w * N + i];
    __syncthreads();
    for (int k = tid; k < H; k += NT) {
        const T *x = sm.x;
        int m0 = 2 * k, m1 = M - 1 - 2 * k;
        T u0 = m0 < H ? -x[3 * H - 1 - m0] - x[3 * H + m0] : x[m0 - H] - x[2 * H - 1 - (m0 - H)];
        T u1 = m1 < H ? -x[3 * H - 1 - m1] - x[3 * H + m1] : x[m1 - H] - x[2 * H - 1 - (m1 - H)];
        sm.W[k] = cmul(mk2<T>(u0, u1), tb.mdct_pre[k]);
    }
    __syncthreads();
    fft_dif<T, LOGM - 1, NT>(sm.W, 1, H + 2, tb.tw, 2);
    for (int k = tid; k < H; k += NT) {
        T2 y = cmul(sm.W[fft_pos<LOGM - 1>(k)], tb.mdct_post[k]);
        out[w * M + 2 * k] = (double)(((T)2 / (T)N) * y.x);
        out[w * M + M - 1 - 2 * k] = (double)(-((T)2 / (T)N) * y.y);
    }
}

// mdct.py:73-88: IMDCT(data, M, M) = unfolded DCT-IV
template <typename T, int LOGM>
__global__ void k_imdct(const double *in, double *out, int n, DevTables<T> tb) {
    using SS = MdctSmem<T, LOGM>;
    using T2 = typename Vec2<T>::type;
    constexpr int M = SS::M, N = 2 * M, H = M / 2, NT = (M / 4 > 0 ? M / 4 : 1);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SS &sm = *reinterpret_cast<SS *>(smem_raw);
    const int tid = threadIdx.x;
    const int64_t w = blockIdx.x;
    for (int i = tid; i < M; i += NT) sm.x[i] = (T)in[w * M + i];
    __syncthreads();
    for (int k = tid; k < H; k += NT) sm.W[k] = cmul(mk2<T>(sm.x[2 * k], sm.x[M - 1 - 2 * k]), tb.mdct_pre[k]);
    __syncthreads();
    fft_dif<T, LOGM - 1, NT>(sm.W, 1, H + 2, tb.tw, 2);
    for (int k = tid; k < H; k += NT) {
        T2 y = cmul(sm.W[fft_pos<LOGM - 1>(k)], tb.mdct_post[k]);
        sm.x[2 * k] = y.x;
        sm.x[M - 1 - 2 * k] = -y.y;
    }
    __syncthreads();
    for (int nn = tid; nn < N; nn += NT) {
        T v = nn < H ? sm.x[nn + H] : (nn < 3 * H ? -sm.x[3 * H - 1 - nn] : -sm.x[nn - 3 * H]);
        out[w * N + nn] = (double)((T)2 * v);
    }
}

// bitalloc.py:129-184, one warp per problem
__global__ void k_bitalloc(int n, const double *bitBudget, const long long *extraBits, int maxMantBits, const double *smr,
                           const uint32_t *lrms, int32_t *bits, long long *diff, BandInfo bands) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n) return;
    const int NB = bands.nBands;
    double s = lane < NB ? smr[(int64_t)p * NB + lane] : 0.0;
    long long d;
    const int nl = lane < NB ? bands.lo[lane + 1] - bands.lo[lane] : 0;
    const long long total0 = (long long)(bitBudget[p] + (double)extraBits[p]);               // int() truncation, bitalloc.py:159
    int b = warp_bitalloc_jump<double>(total0, extraBits[p], maxMantBits, NB, s, lrms[p], nl, &d);
    if (lane < NB) bits[(int64_t)p * NB + lane] = b;
    if (lane == 0) diff[p] = d;
}

// The allocators HEAD does not call (bitalloc.py:22-125), one warp per problem, lane b owns band b.
// mode 0 BitAllocUniform (:22-57), 1 BitAllocConstSNR (:60-91, level = peakSPL per band), 2 BitAllocConstMNR (:94-125, level = SMR).
// status[p] = 1 where the reference's `while remaining_bits > 0` would never end (bits left but no band can take one).
__global__ void k_bitalloc_alt(int n, int mode, const double *bitBudget, int maxMantBits, const double *level, int32_t *bits,
                               int32_t *status, BandInfo bands) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= n) return;
    const int NB = bands.nBands;
    const bool inband = lane < NB;
    const int nl = inband ? bands.lo[lane + 1] - bands.lo[lane] : 0;
    const double budget = bitBudget[p];
    int b = 0, st = 0;
    if (mode == 0) {
        const int total = bands.lo[NB];
        const int per = (int)(budget / (double)total);                       // :32
        // sum(allocation * nLines) is an exact integer sum; the subtraction from bitBudget is the one rounding (:37)
        double remaining = budget - (double)((long long)per * total);
        b = per;
        if (remaining != 0.0) {
            // the walk over bands is sequential but tiny: every lane replays it and keeps its own band's count
            long long line = 0;
            if (total == 0) st = 1;
            while (!st && remaining > 0) {
                const int bd = (int)(line % NB);
                remaining -= (double)(bands.lo[bd + 1] - bands.lo[bd]);
                if (remaining < 0) break;
                if (bd == lane && b < maxMantBits) b += 1;
                line++;
            }
        }
    } else {
        double fl = inband ? level[(int64_t)p * NB + lane] : 0.0;
        double remaining = budget;
        while (remaining > 0) {
            const bool can = inband && b < maxMantBits && remaining - (double)nl >= 0;
            if (__ballot_sync(0xffffffffu, can) == 0u) { st = 1; break; }
            // argmax of the noise floors, first index wins (np.argmax)
            const unsigned long long key = inband ? sortable(fl) : 0ull;
            const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
            const unsigned mh = __reduce_max_sync(0xffffffffu, inband ? hi : 0u);
            const bool c1 = inband && hi == mh;
            const unsigned ml = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
            const int iMax = __ffs(__ballot_sync(0xffffffffu, c1 && lo == ml)) - 1;
            const int nlm = bands.lo[iMax + 1] - bands.lo[iMax];
            const int bm = __shfl_sync(0xffffffffu, b, iMax);
            const bool take = bm < maxMantBits && remaining - (double)nlm >= 0;
            if (take) remaining -= (double)nlm;
            if (lane == iMax) { if (take) b += 1; fl -= 6.0; }
        }
    }
    if (b < 2) b = 0;                                                          // mid-tread: no 1-bit mantissas
    if (b > maxMantBits) b = maxMantBits;
    if (inband) bits[(int64_t)p * NB + lane] = b;
    if (lane == 0) status[p] = st;
}

// Histogram.generateStatistics (Huffman.py:71-83) over n unsigned mantissa codes: occurrences per code and the position of
// each code's first occurrence (the reference's dict keeps first-insertion order, which decides ties when the trainer sorts
// by frequency, :93-108).  Small codes dominate (they are the Huffman-coded ones), so each CTA counts codes < kHistLocal in
// shared memory and merges once; larger codes go straight to global atomics.  Grid-stride, coalesced 16-byte loads.
constexpr int kHistLocal = 4096;
__global__ void __launch_bounds__(256)
k_histogram(const uint32_t *codes, long long n, long long base, int nbins, unsigned long long *counts, unsigned long long *first) {
    __shared__ unsigned cnt[kHistLocal];
    __shared__ unsigned fst[kHistLocal];          // first position seen by this CTA, relative to the CTA's first element + 1 (0 = none)
    for (int i = threadIdx.x; i < kHistLocal; i += blockDim.x) { cnt[i] = 0; fst[i] = 0xffffffffu; }
    __syncthreads();
    // each CTA owns one contiguous span so that 32-bit relative positions are enough for the shared-memory minima
    const long long per = (n + gridDim.x - 1) / gridDim.x;
    const long long lo = per * blockIdx.x, hi = lo + per < n ? lo + per : n;
    for (long long i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const uint32_t c = codes[i];
        if (c < (uint32_t)kHistLocal) {
            atomicAdd(&cnt[c], 1u);
            atomicMin(&fst[c], (unsigned)(i - lo));
        } else if (c < (uint32_t)nbins) {
            atomicAdd(&counts[c], 1ull);
            atomicMin(&first[c], (unsigned long long)(base + i));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHistLocal && i < nbins; i += blockDim.x) {
        if (cnt[i]) {
            atomicAdd(&counts[i], (unsigned long long)cnt[i]);
            atomicMin(&first[i], (unsigned long long)(base + lo + fst[i]));
        }
    }
}

// Huffman.encodeData (Huffman.py:274-309) on caller-supplied unsigned mantissas: total code length under each of
// the 10 tables (escape = escape code + bitAlloc raw bits), strictly-shortest wins, ties keep the lowest ID.
__global__ void k_huff_select(const uint32_t *mag, const int32_t *ba, int n, const unsigned long long *lenLut, EncConsts ec,
                              int32_t *tableID, long long *totals /*[kNTables]*/) {
    __shared__ unsigned long long tot[kNTables];
    if (threadIdx.x < kNTables) tot[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long loc[kNTables];
#pragma unroll
    for (int t = 0; t < kNTables; t++) loc[t] = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        unsigned m = mag[i];
        unsigned long long lw = m < (unsigned)kLenLutSize ? lenLut[m] : 0ull;
#pragma unroll
        for (int t = 0; t < kNTables; t++) {
            unsigned l = (unsigned)(lw >> (5 * t)) & 31u;
            loc[t] += l ? l : (unsigned)(ec.esc_len[t] + ba[i]);
        }
    }
#pragma unroll
    for (int t = 0; t < kNTables; t++) atomicAdd(&tot[t], loc[t]);
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = 1;
        unsigned long long bv = tot[0];
        for (int t = 1; t < kNTables; t++) if (tot[t] < bv) { bv = tot[t]; best = t + 1; }
        *tableID = best;
        for (int t = 0; t < kNTables; t++) totals[t] = (long long)tot[t];
    }
}

// quantize.py
__global__ void k_scale_factor(const double *x, int n, int nScaleBits, int nMantBits, int32_t *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = scale_factor(x[i], nScaleBits, nMantBits);
}
__global__ void k_vquantize_uniform(const double *x, int n, int nBits, unsigned long long *q) {   // :91-117
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long v = quant_mag(fabs(x[i]), nBits);
    if (signbit(x[i])) v += 1ull << (nBits - 1);
    q[i] = v;
}
__global__ void k_vdequantize_uniform(const unsigned long long *q, int n, int nBits, double *x) {   // :120-145
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long mask = 1ull << (nBits - 1), v = q[i];
    bool neg = (v & mask) == mask;
    if (neg) v -= mask;
    double a = 2.0 * (double)v / ((double)(mask << 1) - 1.0);
    x[i] = neg ? -a : a;
}
__global__ void k_vmantissa(const double *x, int n, int scale, int nScaleBits, int nMantBits, unsigned long long *m) {   // :315-342
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (nScaleBits < 0) nScaleBits = 0;
    int largestScale = (1 << nScaleBits) - 1;
    int R = nMantBits + largestScale;
    unsigned long long q = quant_mag(fabs(x[i]), R);
    unsigned long long v = (q << (scale + 1)) >> (R - nMantBits + 1);
    if (signbit(x[i])) v += 1ull << (nMantBits - 1);
    m[i] = v;
}
__global__ void k_vdequantize(int scale, const long long *m, int n, int nScaleBits, int nMantBits, double *x) {   // :345-376
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (nScaleBits < 0) nScaleBits = 0;
    x[i] = dequant(scale, m[i], (1 << nScaleBits) - 1, nMantBits);
}

}  // namespace pac
