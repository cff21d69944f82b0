// mdct.cuh -- K1 as its own kernel (fp32 fast mode): PCM -> signed fractions (pcmfile.py:66-100, quantize.py:120-145),
// SineWindow (window.py:27-39), MDCT (mdct.py:49-71), overall scale (codec.py:237-246), for one stereo block per CTA.
//
// Why not inside k_analysis (where it lived in round 1): an fp32 FFT leaves an error floor of ~1e-7 x (largest line) on EVERY
// line, far more than north_star's 1e-5 relative on the weak lines, so the transform runs in fp64 -- and as a section of the
// 64-register, 4-CTA/SM analysis kernel that fp64 code ran at 31 ns per block (6 % of the HBM roofline of its 12.3 KB).  On its
// own it is a memory-shaped kernel: 16-byte PCM loads, TDAC fold with the window applied on the fly, a register-blocked
// radix-8 x 8 x 8 (8 x 8 x 4 for 256 points) FFT of the M/2 folded complex points per channel with the last pass autosorting,
// post-twiddle, block maximum -> overall scale, coalesced fp32 stores of the scaled lines.  k_analysis then reads those lines
// back (8 KB per block) and overwrites them in place with the LRMS-selected ones.
// Algebra as tests/model_analysis.py:mdct_fold_fft (DCT-IV by an M/2-point complex FFT).
#pragma once
#include "common.cuh"
#include "fft.cuh"

namespace pac {

struct MdctArgs {
    // ---- input: exactly one of pcm / blocks
    const int16_t *pcm;          // [S][strideSamples][2]
    int64_t strideSamples;
    const int64_t *nSamples;     // [S]
    const double *blocks;        // [nwork][2][N] raw signed fractions (per-block API)
    int S, b0, nb;               // work item w: stream s = w / nb, block b = b0 + w % nb
    int64_t nwork;
    int nScaleBits;
    float *lines;                // [nwork][2][M] scaled L/R lines (2^overallScale * MDCT, rounded once to fp32)
    uint8_t *oscale;             // [nwork][2]
    DevTables<double> tabd;
};

template <int LOGM>
struct EncMdctSmem {
    static constexpr int M = 1 << LOGM, H = M / 2;
    double2 Z[2][H + H / 8];     // folded points / FFT work per channel (padded)
    union {                      // the window's samples are dead once folded; the same 8 KB then stage the lines
        int pcm[2 * M];          // the block's 2048 stereo frames (int16 pairs)
        float v[2][M];           // unscaled lines, staged for the coalesced store
    };
    double2 twH[H];              // exp(-2 pi i m / H): the first pass's twiddles (fetched per CTA once; as global loads they missed the
    double2 twL[H / 8];          // small L1 this kernel leaves and sat on the L2 latency, ncu); exp(-2 pi i m / (H/8)) for the second pass
    double red[2][4];
    int oscale[2];
};

// CTA = M/8 threads: the first half works on channel 0 in the FFT passes, the second on channel 1
template <int LOGM, bool PCM>
__global__ void __launch_bounds__((1 << LOGM) / 8)
k_mdct_enc(const __grid_constant__ MdctArgs a) {
    using SM = EncMdctSmem<LOGM>;
    constexpr int M = 1 << LOGM, N = 2 * M, H = M / 2, NT = M / 8, HT = NT / 2;    // HT threads per channel = H/8 radix-8 items
    constexpr int LOGH = LOGM - 1;
    constexpr int R3 = (LOGH % 3 == 0) ? 8 : 4;                                  // 512 = 8*8*8, 256 = 8*8*4
    static_assert(LOGH == 9 || LOGH == 8, "k_mdct_enc is written for 1024 / 512 lines");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM &sm = *reinterpret_cast<SM *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevTables<double> &td = a.tabd;
    const int ch = tid / HT, t = tid - ch * HT;              // FFT role
    for (int m = tid; m < H; m += NT) sm.twH[m] = td.tw[2 * m];
    for (int m = tid; m < H / 8; m += NT) sm.twL[m] = td.tw[16 * m];
    // the window [(b-1) M, (b+1) M) of work item w, zero outside [0, n): 16-byte loads (4 stereo frames) into registers.  The NEXT
    // item's window is requested before the current one is transformed, so its HBM latency hides behind the transform.
    constexpr int NV = N / 4 / NT;                          // int4 per thread
    int4 win[NV];
    auto fetch = [&](int64_t w) -> bool {
        const int s = a.nwork <= 0xffffffffll ? (int)((uint32_t)w / (uint32_t)a.nb) : (int)(w / a.nb);
        const int b = a.b0 + (int)(w - (int64_t)s * a.nb);
        const int64_t ns = a.nSamples[s];
        const int64_t nblk = (ns + M - 1) / M + 1;
        if (b >= nblk) return false;                         // uniform for the CTA
        const int *p32 = reinterpret_cast<const int *>(a.pcm) + (int64_t)s * a.strideSamples;
        const int64_t base = (int64_t)(b - 1) * M;
        const bool vec = (reinterpret_cast<uintptr_t>(p32) & 15) == 0;        // row base 16-byte aligned
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const int64_t si = base + 4 * (tid + NT * i);    // base is a multiple of M, so 4-frame groups are aligned
            int4 v4 = make_int4(0, 0, 0, 0);
            if (si >= 0 && si + 4 <= ns && vec) v4 = __ldg(reinterpret_cast<const int4 *>(p32 + si));
            else if (si + 4 > 0 && si < ns) {
                int r[4];
#pragma unroll
                for (int j = 0; j < 4; j++) r[j] = (si + j >= 0 && si + j < ns) ? __ldg(p32 + si + j) : 0;
                v4 = make_int4(r[0], r[1], r[2], r[3]);
            }
            win[i] = v4;
        }
        return true;
    };
    bool have = false;
    if (PCM && (int64_t)blockIdx.x < a.nwork) have = fetch(blockIdx.x);
    __syncthreads();
    for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
        if (PCM) {
            const bool cur = have;
            if (cur) {
#pragma unroll
                for (int i = 0; i < NV; i++) *reinterpret_cast<int4 *>(&sm.pcm[4 * (tid + NT * i)]) = win[i];
            }
            have = (w + gridDim.x < a.nwork) ? fetch(w + gridDim.x) : false;
            if (!cur) continue;                              // a block past the end of its stream (uniform for the CTA)
            __syncthreads();
        }
        // windowed sample n of channel c as the reference forms it: x = +-2|code|/65535 (quantize.py:141; -32768 -> 0), times sin
        auto xw = [&](int c, int n) -> double {
            double v;
            if (PCM) {
                const int pr = sm.pcm[n];
                const int sv = c ? (pr >> 16) : (int)(short)(pr & 0xffff);
                int code = sv < 0 ? -sv : sv;
                if (code & 32768) code -= 32768;
                v = (double)(sv < 0 ? -code : code) * (2.0 / 65535.0);
            } else v = a.blocks[(w * 2 + c) * N + n];
            return v * td.sinw[n];
        };
        // ---- fold (TDAC) to H complex points per channel, pre-twiddle
        for (int n = tid; n < H; n += NT) {
            const int m0 = 2 * n, m1 = M - 1 - 2 * n;
            const double2 pre = td.mdct_pre[n];
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const double u0 = m0 < H ? -xw(c, 3 * H - 1 - m0) - xw(c, 3 * H + m0) : xw(c, m0 - H) - xw(c, 2 * H - 1 - (m0 - H));
                const double u1 = m1 < H ? -xw(c, 3 * H - 1 - m1) - xw(c, 3 * H + m1) : xw(c, m1 - H) - xw(c, 2 * H - 1 - (m1 - H));
                sm.Z[c][pad8(n)] = cmul(mk2<double>(u0, u1), pre);
            }
        }
        __syncthreads();
        // ---- H-point FFT per channel, radix 8 x 8 x R3, in place; tw[m * (M / H) ...]: td.tw[m] = exp(-2 pi i m / M), so exp(-2 pi i m / H) = tw[2 m]
        double2 *Z = sm.Z[ch];
        {   // pass 1: elements j + L r, L = H/8; twiddle W_H^(j p)
            constexpr int L = H / 8;
            const int j = t;                                 // HT == L
            double2 x[8];
#pragma unroll
            for (int r = 0; r < 8; r++) x[r] = Z[pad8(j + L * r)];
            dft8(x);
#pragma unroll
            for (int p = 0; p < 8; p++) {
                double2 y = x[brev3(p)];
                if (p) y = cmul(y, sm.twH[j * p]);
                Z[pad8(j + L * p)] = y;
            }
        }
        __syncthreads();
        {   // pass 2: sub-transforms of length L1 = H/8: elements L1 g + j + L2 r, L2 = L1/8; twiddle W_L1^(j p)
            constexpr int L1 = H / 8, L2 = L1 / 8;
            const int g = t / L2, j = t - g * L2;
            double2 x[8];
#pragma unroll
            for (int r = 0; r < 8; r++) x[r] = Z[pad8(L1 * g + j + L2 * r)];
            dft8(x);
#pragma unroll
            for (int p = 0; p < 8; p++) {
                double2 y = x[brev3(p)];
                if (p) y = cmul(y, sm.twL[j * p]);
                Z[pad8(L1 * g + j + L2 * p)] = y;
            }
        }
        __syncthreads();
        {   // pass 3: radix R3 on consecutive elements, autosorted: frequency k = p1 + 8 p2 + 64 p3 (position (H/8) p1 + (H/64) p2 + p3)
            constexpr int L2 = H / 64;                       // 8 (H = 512) or 4 (H = 256) = R3
            static_assert(L2 == R3, "last pass covers one whole sub-transform");
            double2 x[8];
            // thread t owns the sub-transform at positions R3 t .. R3 t + R3 - 1 (HT * R3 == H when R3 == 8; for R3 == 4 two per thread)
            constexpr int PER = H / (HT * R3);               // 1 or 2
            double2 y[PER][8];
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int g = t + HT * u;                    // sub-transform index: p1 = g / 8, p2 = g % 8
#pragma unroll
                for (int r = 0; r < R3; r++) x[r] = Z[pad8(R3 * g + r)];
                if (R3 == 8) {
                    dft8(x);
#pragma unroll
                    for (int p = 0; p < 8; p++) y[u][p] = x[brev3(p)];
                } else {
                    radix2(x[0], x[2]); radix2(x[1], x[3]);
                    x[3] = mk2<double>(x[3].y, -x[3].x);
                    radix2(x[0], x[1]); radix2(x[2], x[3]);
                    y[u][0] = x[0]; y[u][1] = x[2]; y[u][2] = x[1]; y[u][3] = x[3];
                }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int g = t + HT * u;
                const int kb = (g >> 3) + 8 * (g & 7);
#pragma unroll
                for (int p = 0; p < R3; p++) Z[pad8(kb + 64 * p)] = y[u][p];
            }
        }
        __syncthreads();
        // ---- post-twiddle: X[2k] = (2/N) Re(y_k e^{-i pi (k + 1/4)/M}), X[M-1-2k] = -(2/N) Im(...); block maximum per channel
        double mx = 0.0;
        for (int k = t; k < H; k += HT) {
            const double2 y = cmul(Z[pad8(k)], td.mdct_post[k]);
            const double v0 = (2.0 / (double)N) * y.x, v1 = -(2.0 / (double)N) * y.y;
            sm.v[ch][2 * k] = (float)v0;                     // rounding to fp32 commutes with the power-of-two overall scale
            sm.v[ch][M - 1 - 2 * k] = (float)v1;
            mx = fmax(mx, fmax(fabs(v0), fabs(v1)));
        }
        mx = warp_max(mx);
        if (lane == 0) sm.red[0][warp] = mx;                 // warps 0 .. NT/64-1 belong to channel 0, the rest to channel 1
        __syncthreads();
        if (tid < 2) {
            constexpr int WPC = NT / 64;                     // warps per channel
            double m = 0.0;
            for (int i = 0; i < WPC; i++) m = fmax(m, sm.red[0][tid * WPC + i]);
            // the overall scale sees the fp32-rounded maximum, exactly as the fused kernel of round 1 did (codec.py:245, default nMantBits = 5)
            sm.oscale[tid] = scale_factor((double)(float)m, a.nScaleBits, 5);
        }
        __syncthreads();
        {
            const float s0 = (float)(1 << sm.oscale[0]), s1 = (float)(1 << sm.oscale[1]);
            float4 *dst = reinterpret_cast<float4 *>(a.lines + w * 2 * M);
            const float4 *src = reinterpret_cast<const float4 *>(&sm.v[0][0]);
            for (int q = tid; q < 2 * M / 4; q += NT) {
                float4 v4 = src[q];
                const float sc = q < M / 4 ? s0 : s1;        // codec.py:246
                v4.x *= sc; v4.y *= sc; v4.z *= sc; v4.w *= sc;
                dst[q] = v4;
            }
            if (tid < 2) a.oscale[w * 2 + tid] = (uint8_t)sm.oscale[tid];
        }
        __syncthreads();
    }
}

}  // namespace pac
