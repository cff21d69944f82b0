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

// X[k] (0 <= k <= M) of a real sequence whose packed (even + i*odd) M-point FFT sits in Z (digit-reversed).
template <typename T, int LOGM, bool PAD = false>
__device__ __forceinline__ typename Vec2<T>::type rfft_split(const typename Vec2<T>::type *Z, int k,
                                                             const typename Vec2<T>::type *__restrict__ tw_split) {
    using T2 = typename Vec2<T>::type;
    constexpr int Mm = (1 << LOGM) - 1;
    T2 zk = Z[fft_pad<PAD>(fft_pos<LOGM>(k & Mm))];
    T2 zm = cconj(Z[fft_pad<PAD>(fft_pos<LOGM>(((1 << LOGM) - k) & Mm))]);
    T2 e = cadd(zk, zm), d = csub(zk, zm);
    e.x *= (T)0.5; e.y *= (T)0.5;
    T2 o = mk2<T>(d.y * (T)0.5, -d.x * (T)0.5);      // (zk - conj(zm)) / (2i)
    return cadd(e, cmul(tw_split[k], o));
}

}  // namespace pac
