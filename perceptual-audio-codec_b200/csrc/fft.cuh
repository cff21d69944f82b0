// fft.cuh -- shared-memory complex FFT (in-place decimation-in-frequency, radix-4 stages + one radix-2
// stage when log2(N) is odd), batched across the CTA.  Output is left in digit-reversed order; readers
// use fft_pos<LOGN>(k) to find frequency k.  Index math proven in tests/test_model.py-style numpy
// prototypes (see DESIGN.md "FFT").
#pragma once
#include "common.cuh"

namespace pac {

// position of frequency k after fft_dif<LOGN>
template <int LOGN>
__device__ __forceinline__ int fft_pos(int k) {
    int p = 0;
    int L = 1 << LOGN;
#pragma unroll
    for (int s = 0; s < LOGN / 2; s++) {
        int q = L >> 2;
        p += (k & 3) * q;
        k >>= 2;
        L >>= 2;
    }
    if (LOGN & 1) p += k;
    return p;
}

// Padded layout (PAD = true): logical element i lives at i + (i >> 4), i.e. one unused element after every 16.  With
// 8-byte elements that turns the bank-aliased strides of the late stages (16-element groups 128 B apart) and of the
// digit-reversed read-out into conflict-free or 2-4-way accesses (ncu: 43 % of the kernel's shared-memory wavefronts
// were conflict replays, 2/3 of them here).  A row then needs N + N/16 elements.
template <bool PAD>
__device__ __forceinline__ int fft_pad(int i) { return PAD ? i + (i >> 4) : i; }

// buf: nbatch transforms of N = 2^LOGN points, transform b at buf + b*bstride.
// tw[m*twstride] = exp(-2 pi i m / N).  All NTHREADS threads of the CTA must call.  Ends with a barrier.
template <typename T, int LOGN, int NTHREADS, bool PAD = false>
__device__ __forceinline__ void fft_dif(typename Vec2<T>::type *buf, int nbatch, int bstride,
                                        const typename Vec2<T>::type *__restrict__ tw, int twstride) {
    using T2 = typename Vec2<T>::type;
    constexpr int N = 1 << LOGN;
    constexpr int NQ = N / 4 > 0 ? N / 4 : 1;
    const int tid = threadIdx.x;
    const int total = nbatch * NQ;
    int logL = LOGN;
#pragma unroll
    for (int s = 0; s < LOGN / 2; s++, logL -= 2) {
        const int L = 1 << logL;
        const int q = L >> 2;
        const int ts = (N >> logL) * twstride;
        for (int w = tid; w < total; w += NTHREADS) {
            int bt = w / NQ, r = w - bt * NQ;
            int g = r >> (logL - 2), j = r & (q - 1);
            // the four legs are q apart: whole 16-groups when q >= 16, inside one group (no pad between them) when q < 16
            const int qs = (PAD && q >= 16) ? q + (q >> 4) : q;
            T2 *p = buf + bt * bstride + fft_pad<PAD>((g << logL) + j);
            T2 a0 = p[0], a1 = p[qs], a2 = p[2 * qs], a3 = p[3 * qs];
            T2 b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3), d = csub(a1, a3);
            T2 b3 = mk2<T>(d.y, -d.x);                 // -i * (a1 - a3)
            T2 y0 = cadd(b0, b2), y2 = csub(b0, b2), y1 = cadd(b1, b3), y3 = csub(b1, b3);
            if (q > 1) {
                T2 w1 = tw[j * ts], w2 = tw[2 * j * ts], w3 = tw[3 * j * ts];
                y1 = cmul(y1, w1); y2 = cmul(y2, w2); y3 = cmul(y3, w3);
            }
            p[0] = y0; p[qs] = y1; p[2 * qs] = y2; p[3 * qs] = y3;
        }
        __syncthreads();
    }
    if (LOGN & 1) {
        const int tot2 = nbatch * (N / 2);
        for (int w = tid; w < tot2; w += NTHREADS) {
            int bt = w / (N / 2), r = w - bt * (N / 2);
            T2 *p = buf + bt * bstride + fft_pad<PAD>(2 * r);
            T2 a0 = p[0], a1 = p[1];
            p[0] = cadd(a0, a1); p[1] = csub(a0, a1);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Register-blocked 1024-point FFT (fp32 batch of the analysis kernel): radix 16 x 16 x 4, in place, decimation in
// frequency.  A thread holds 16 points in registers and does a whole 16-point DFT (two levels of radix-4 butterflies with the
// W16 twiddles as constants) between one shared-memory read and one write: three passes through shared memory and three
// barriers instead of the five of the radix-4 version, and the index/twiddle-address arithmetic is paid per 16 points
// instead of per 4 (ncu, round 1: the FFT butterflies + complex helpers were 17.7 % of the analysis kernel's instructions).
//   pass 1: elements j + 64 r            (j < 64)            twiddle W_1024^(j p)
//   pass 2: elements 64 g + j + 4 r      (g < 16, j < 4)     twiddle W_64^(j p)
//   pass 3: radix 4 on 4 consecutive elements, no twiddle
// The last pass writes frequency k = p1 + 16 p2 + 256 p3 (which its in-place form would leave at 64 p1 + 4 p2 + p3) to its
// natural position.  With the padded layout (one spare element per 16) every pass is conflict-free for 64-bit accesses.
// ------------------------------------------------------------------------------------------------------------------
// position of frequency k: the register-blocked 1024-point FFT leaves natural order, the radix-4 one digit-reversed order
template <int LOGN, bool R16>
__device__ __forceinline__ int fft_pos_sel(int k) {
    if constexpr (R16 && LOGN == 10) return k;
    else return fft_pos<LOGN>(k);
}

__device__ __forceinline__ void radix4_f(float2 &a0, float2 &a1, float2 &a2, float2 &a3) {
    const float2 b0 = cadd(a0, a2), b1 = csub(a0, a2), b2 = cadd(a1, a3), d = csub(a1, a3);
    const float2 b3 = make_float2(d.y, -d.x);                  // -i (a1 - a3)
    a0 = cadd(b0, b2); a2 = csub(b0, b2); a1 = cadd(b1, b3); a3 = csub(b1, b3);
}

// in-register 16-point DFT: y[d + 4c] (natural order) from x[a + 4b]
__device__ __forceinline__ void dft16_f(float2 (&x)[16]) {
    const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, H = 0.70710678118654752f;
#pragma unroll
    for (int a = 0; a < 4; a++) radix4_f(x[a], x[a + 4], x[a + 8], x[a + 12]);     // t[a][d] sits in x[a + 4 d]
    // twiddles W16^(a d), a, d = 1..3
    x[1 + 4] = cmul(x[1 + 4], make_float2(C1, -S1));           // a=1,d=1: W^1
    x[1 + 8] = cmul(x[1 + 8], make_float2(H, -H));             // a=1,d=2: W^2
    x[1 + 12] = cmul(x[1 + 12], make_float2(S1, -C1));         // a=1,d=3: W^3
    x[2 + 4] = cmul(x[2 + 4], make_float2(H, -H));             // a=2,d=1: W^2
    x[2 + 8] = make_float2(x[2 + 8].y, -x[2 + 8].x);           // a=2,d=2: W^4 = -i
    x[2 + 12] = cmul(x[2 + 12], make_float2(-H, -H));          // a=2,d=3: W^6
    x[3 + 4] = cmul(x[3 + 4], make_float2(S1, -C1));           // a=3,d=1: W^3
    x[3 + 8] = cmul(x[3 + 8], make_float2(-H, -H));            // a=3,d=2: W^6
    x[3 + 12] = cmul(x[3 + 12], make_float2(-C1, S1));         // a=3,d=3: W^9
#pragma unroll
    for (int d = 0; d < 4; d++) radix4_f(x[4 * d], x[4 * d + 1], x[4 * d + 2], x[4 * d + 3]);   // -> y[d + 4c] in x[4 d + c]
}

// buf: nbatch transforms of 1024 points (padded layout when PAD), transform b at buf + b*bstride.  tw[m] = exp(-2 pi i m / 1024).
template <int NTHREADS, bool PAD>
__device__ __forceinline__ void fft1024_r16(float2 *buf, int nbatch, int bstride, const float2 *__restrict__ tw) {
    const int tid = threadIdx.x;
    // ---- pass 1
    for (int w = tid; w < nbatch * 64; w += NTHREADS) {
        const int bt = w >> 6, j = w & 63;
        float2 *p = buf + bt * bstride;
        float2 x[16];
#pragma unroll
        for (int r = 0; r < 16; r++) x[r] = p[fft_pad<PAD>(j + 64 * r)];
        dft16_f(x);
#pragma unroll
        for (int d = 0; d < 4; d++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int pp = d + 4 * c;                       // output index of x[4 d + c]
                float2 y = x[4 * d + c];
                if (pp) y = cmul(y, tw[j * pp]);
                p[fft_pad<PAD>(j + 64 * pp)] = y;
            }
    }
    __syncthreads();
    // ---- pass 2
    for (int w = tid; w < nbatch * 64; w += NTHREADS) {
        const int bt = w >> 6, g = (w >> 2) & 15, j = w & 3;
        float2 *p = buf + bt * bstride;
        float2 x[16];
#pragma unroll
        for (int r = 0; r < 16; r++) x[r] = p[fft_pad<PAD>(64 * g + j + 4 * r)];
        dft16_f(x);
#pragma unroll
        for (int d = 0; d < 4; d++)
#pragma unroll
            for (int c = 0; c < 4; c++) {
                const int pp = d + 4 * c;
                float2 y = x[4 * d + c];
                if (pp) y = cmul(y, tw[16 * j * pp]);
                p[fft_pad<PAD>(64 * g + j + 4 * pp)] = y;
            }
    }
    __syncthreads();
    // ---- pass 3: radix 4 on consecutive elements, AUTOSORTED: frequency k = p1 + 16 p2 + 256 p3 sits at position 64 p1 + 4 p2 + p3;
    // every thread keeps its groups in registers across a barrier and writes them to their natural places (lanes -> k stride 16
    // -> 17 padded elements: conflict-free), so the readers index by k directly -- no digit-reversed gathers (which cost 4-way bank
    // conflicts on every spectrum read, ncu) and no index arithmetic
    // (two groups per thread are in flight across each barrier: writes of one transform only collide with reads of the same one)
    for (int w0 = 0; w0 < nbatch * 256; w0 += 2 * NTHREADS) {
        float2 y[2][4];
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int w = w0 + tid + i * NTHREADS;
            if (w < nbatch * 256) {
                const int bt = w >> 8, g = w & 255;
                const float2 *p = buf + bt * bstride + fft_pad<PAD>(4 * g);
                y[i][0] = p[0]; y[i][1] = p[1]; y[i][2] = p[2]; y[i][3] = p[3];
                radix4_f(y[i][0], y[i][1], y[i][2], y[i][3]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const int w = w0 + tid + i * NTHREADS;
            if (w < nbatch * 256) {
                const int bt = w >> 8, g = w & 255;
                float2 *p = buf + bt * bstride;
                const int kb = (g >> 4) + 16 * (g & 15);
#pragma unroll
                for (int p3 = 0; p3 < 4; p3++) p[fft_pad<PAD>(kb + 256 * p3)] = y[i][p3];
            }
        }
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------------------------
// Radix-8 building blocks of the register-blocked M/2-point FFTs (encoder k_mdct_enc in fp64, decoder k_synth in fp32).
// pad8: one spare element per 8 -- the radix-8 passes, the autosorting write and the natural-order reads are conflict-free.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pad8(int i) { return i + (i >> 3); }

template <typename V> __device__ __forceinline__ void radix2(V &a, V &b) { const V t = csub(a, b); a = cadd(a, b); b = t; }

// in-register 8-point DFT (decimation in frequency), output y[p] left in x[brev3(p)]
template <typename V>
__device__ __forceinline__ void dft8(V (&x)[8]) {
    using T = decltype(x[0].x);
    const T H = (T)0.70710678118654752440;
#pragma unroll
    for (int i = 0; i < 4; i++) radix2(x[i], x[i + 4]);
    // twiddles W8^i on the lower half: 1, (1-i)/sqrt2, -i, (-1-i)/sqrt2
    { const V t = x[5]; x[5].x = (t.x + t.y) * H; x[5].y = (t.y - t.x) * H; }
    { const V t = x[6]; x[6].x = t.y; x[6].y = -t.x; }
    { const V t = x[7]; x[7].x = (t.y - t.x) * H; x[7].y = -(t.x + t.y) * H; }
#pragma unroll
    for (int h = 0; h < 8; h += 4) {
        radix2(x[h], x[h + 2]); radix2(x[h + 1], x[h + 3]);
        { const V t = x[h + 3]; x[h + 3].x = t.y; x[h + 3].y = -t.x; }       // W4^1 = -i
        radix2(x[h], x[h + 1]); radix2(x[h + 2], x[h + 3]);
    }
}
__device__ __forceinline__ int brev3(int p) { return ((p & 1) << 2) | (p & 2) | (p >> 2); }

// X[k] (0 <= k <= M) of a real sequence whose packed (even + i*odd) M-point FFT sits in Z (digit-reversed).
template <typename T, int LOGM, bool PAD = false, bool R16 = false>
__device__ __forceinline__ typename Vec2<T>::type rfft_split(const typename Vec2<T>::type *Z, int k,
                                                             const typename Vec2<T>::type *__restrict__ tw_split) {
    using T2 = typename Vec2<T>::type;
    constexpr int Mm = (1 << LOGM) - 1;
    T2 zk = Z[fft_pad<PAD>(fft_pos_sel<LOGM, R16>(k & Mm))];
    T2 zm = cconj(Z[fft_pad<PAD>(fft_pos_sel<LOGM, R16>(((1 << LOGM) - k) & Mm))]);
    T2 e = cadd(zk, zm), d = csub(zk, zm);
    e.x *= (T)0.5; e.y *= (T)0.5;
    T2 o = mk2<T>(d.y * (T)0.5, -d.x * (T)0.5);      // (zk - conj(zm)) / (2i)
    return cadd(e, cmul(tw_split[k], o));
}

}  // namespace pac
