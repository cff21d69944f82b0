// analysis.cuh -- K1+K2+K3 fused: one CTA per stereo block does
//   PCM -> signed fractions (pcmfile.py:66-100, quantize.py:120-145)
//   raw-FFT M/S decision per band (codec.py:96-102)
//   SineWindow + MDCT + overall scale (codec.py:237-246, window.py:27-39, mdct.py:49-71, quantize.py:148-177)
//   getStereoMaskThreshold (psychoac.py:506-682): six masked-threshold curves (calcBTHR :409-456,
//   findpeaks :158-191, Masker :66-120), MLD (:349-372), per-band max SMR (:458-504), LRMS select (:662-682)
// entirely out of shared memory / registers; HBM traffic is the algorithmic 4 KB in + ~8.6 KB out per block.
//
// The algebra (real FFTs via packed complex FFTs, MDCT via fold + M/2-point FFT, Hann^p windows of the
// M/S signals as 3-tap frequency-domain convolutions) is stated in numpy in tests/model_analysis.py and
// checked there against dumps of the reference.
#pragma once
#include "common.cuh"
#include <type_traits>

#include "fft.cuh"

#ifndef PAC_ANALYSIS_CTAS
#define PAC_ANALYSIS_CTAS 4      // resident fp32 analysis CTAs per SM the kernel is compiled for (64 registers per thread)
#endif

namespace pac {

__device__ __forceinline__ void cp_async16_a(void *smemDst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smemDst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_a() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_a() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <typename T>
struct AnalysisArgs {
    // ---- input: exactly one of pcm / blocks
    const int16_t *pcm;          // [S][strideSamples][2]
    int64_t strideSamples;
    const int64_t *nSamples;     // [S]
    const double *blocks;        // [nwork][2][N] raw signed fractions
    int S, b0, nb;               // work item w: stream s = w / nb, block b = b0 + w % nb
    int64_t nwork;
    int nScaleBits;
    // ---- outputs, indexed by w.  fp32 mode: lines / oscale are ALSO inputs -- k_mdct (mdct.cuh) has left the scaled L/R lines and
    //      the overall scales there; the lines are overwritten in place with the LRMS-selected ones
    T *lines;                    // [nwork][2][M]  LRMS-selected scaled lines
    T *smr;                      // [nwork][2][kMaxBands]
    T *bmax;                     // [nwork][2][kMaxBands] max |selected line| per band
    uint8_t *oscale;             // [nwork][2]
    uint32_t *lrms;              // [nwork]
    T *dbg_mdct;                 // [nwork][2][M]  (optional) scaled L/R lines
    T *dbg_bthr;                 // [nwork][6][M]  (optional) L,R,M,S,M',S' thresholds in dB
    DevTables<T> tab;
    DevTables<double> tabd;      // double tables: fp32 mode computes the MDCT in fp64 and splits Bark values hi+lo
    FastTables ft;               // fp32 mode only
    BandInfo bands;
    uint32_t poisonOn, poison, smemWords;   // debugging aid (PAC_POISON_SMEM): refill shared memory before every block
};

// extra shared memory of the fp32 fast threshold evaluation.  The three dense per-bin arrays of a curve (Sp, V, Wq: 16 400 B)
// are NOT here: they live in whichever spectrum buffer is dead while the curve runs (W during the L/R curves, XF afterwards),
// which is what lets four CTAs share an SM.
template <int LOGM>
struct FastSmem {
    static constexpr int M = 1 << LOGM;
    double wtot[16];
    float4 loud[M / 2 + 1];       // (+1: a sentinel with exponent -inf, the "no masker" operand of the 4-wide survivor loop)
                                  // maskers louder than 40 dB: (B_hi, up, B_lo, first upper-skirt line as int bits) with
                                  // B = c0 - up/2 - up * z_masker formed in double: exponent at line i = up * z_i + B
    unsigned short loudBase[M / 4 + 4];   // number of loud maskers below bin 4t ...
    unsigned char loudFlag[M / 4 + 4];    // ... and which of the bins 4t .. 4t+3 are loud (prefix at any bin = base + popc)
    float totD[16], totA[16];
    float4 stash[2][M / 4];       // per thread: the static records (FastTables::lineRec) of its two lines in the warp's second half-chunk
};
// scratch layout inside the dead spectrum buffer
template <int LOGM>
struct CurveScratch {
    static constexpr int M = 1 << LOGM;
    double Sp[M + 2];             // exclusive prefix sums (in double) of the plateau intensities A_k over the bins
    float SD[M], SA[M];           // per bin: descending scan of all maskers' intensities (lower skirts), ascending scan of the quiet
                                  // maskers' (upper skirts), both with the fixed -27 dB/Bark decay between the bins' Bark positions
};
struct NoSmem {};

template <typename T, int LOGM, bool FASTK = false>
struct AnalysisSmem {
    static constexpr int M = 1 << LOGM;
    static constexpr int N = 2 * M;
    static constexpr int NT = M / 4;
    using T2 = typename Vec2<T>::type;
    // fp32 mode runs its 1024-point FFTs in the padded layout of fft.cuh (one spare element per 16): rows of M + M/16
    static constexpr int ROW = FASTK ? M + M / 16 : M + 2;
    T2 XF[2][ROW];         // time samples x[ch][n] (viewed as T[2][2*ROW], padded in fp32 mode); later F1[ch][0..M]
    T2 W[2][ROW];          // FFT work; later F2_M, F2_S; finally the four per-line SMR candidate arrays
    T Lb[FASTK ? 1 : 2][FASTK ? 4 : M];   // scaled MDCT lines L, R (fp64 / stage kernels).  fp32 mode keeps no buffer for them: they land in
                                          // W by cp.async once the last curve has consumed its spectrum (8 KB less per CTA; a 196 KB carve-out
                                          // with 60 KB of L1 instead of 27 measured no different: 1584.5 vs 1585.0 ms)
    T P[M + 8];            // |spectrum|^2 of the current curve
    static constexpr int MLM = FASTK ? 1 : M / 2;                // the direct evaluation's masker list (fp64 / mono kernels)
    T mz[MLM], mp[MLM], ml[MLM];         // Bark position, SPL, 0.367*max(SPL-40,0)
    T mzlo[MLM];                         // (low part of the Bark position when T = float)
    T red[64];
    int wsum[NT / 32 + 1];
    int cnt;
    uint32_t lrms;
    int oscale[2];
    typename std::conditional<FASTK, FastSmem<LOGM>, NoSmem>::type fs;
};

// spread + accumulate one masker into 4 lines.  fp64 keeps the reference's operation order
// (psychoac.py:111-120); fp32 folds the constants and uses ex2.approx.
__device__ __forceinline__ void spread4(double (&acc)[4], const double (&zl)[4], const double (&)[4], double zm, double,
                                        double pm, double lev, double drop) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
        double dz = zl[j] - zm;
        double a = fabs(dz);
        double spreadv = ((dz >= 0.0 ? lev : 0.0) - 27.0) * (a > 0.5 ? a - 0.5 : 0.0);
        double spl = pm + spreadv - drop;
        acc[j] += exp10((spl - 96.0) / 10.0);
    }
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void spread4(float (&acc)[4], const float (&zl)[4], const float (&zlo)[4], float zm, float zmlo,
                                        float pm, float lev, float drop) {
    const float K = 0.33219280948873623f;           // log2(10)/10
    float c0 = (pm - drop - 96.0f) * K;
    float up = (lev - 27.0f) * K, dn = -27.0f * K;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float dz = (zl[j] - zm) + (zlo[j] - zmlo);  // Bark distance from hi+lo pairs (a single fp32 costs 3e-5 dB)
        float a = fmaxf(fabsf(dz) - 0.5f, 0.0f);
        float sl = dz >= 0.0f ? up : dn;
        acc[j] += ex2_approx(fmaf(sl, a, c0));
    }
}

// DFT of hann*y at bin k (0 <= k < M) from F = DFT(y)[0..M], y real:  .5 F[k] - .25 (w F[k-1] + conj(w) F[k+1])
template <typename T>
__device__ __forceinline__ typename Vec2<T>::type hann_tap(const typename Vec2<T>::type *F, int k,
                                                           typename Vec2<T>::type hw, typename Vec2<T>::type hwc) {
    using T2 = typename Vec2<T>::type;
    T2 fm = k == 0 ? cconj(F[1]) : F[k - 1];
    T2 fp = F[k + 1];
    T2 t = cadd(cmul(hw, fm), cmul(hwc, fp));
    return mk2<T>((T)0.5 * F[k].x - (T)0.25 * t.x, (T)0.5 * F[k].y - (T)0.25 * t.y);
}

// One masked-threshold curve (calcBTHR body after the FFT, psychoac.py:431-456): power spectrum from `spec`,
// findpeaks (:158-191), masker SPLs (:448), spreading over this thread's 4 lines (:447-452), + threshold in quiet,
// -> dB.  All threads of the CTA must call; uses sm.P / sm.mz / sm.mp / sm.ml / sm.wsum / sm.cnt.
template <typename T, int LOGM, class SMEM, class Spec>
__device__ __forceinline__ void masked_curve(SMEM &sm, const DevTables<T> &tb, Spec spec, T drop,
                                             const T (&zl)[4], const T (&zlo)[4], const T (&tiq)[4], const double *zpeakd,
                                             T (&thr)[4]) {
    using T2 = typename Vec2<T>::type;
    constexpr int M = 1 << LOGM, NT = M / 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int k = tid + NT * j;
        T2 v = spec(k);
        sm.P[k] = v.x * v.x + v.y * v.y;
    }
    __syncthreads();
    // findpeaks on 4 consecutive bins per thread, ordered compaction
    int k0 = 4 * tid;
    unsigned flags = 0;
    {
        T pw[6];
        pw[0] = k0 > 0 ? sm.P[k0 - 1] : (T)0;
        pw[1] = sm.P[k0]; pw[2] = sm.P[k0 + 1]; pw[3] = sm.P[k0 + 2]; pw[4] = sm.P[k0 + 3];
        pw[5] = k0 + 4 < M ? sm.P[k0 + 4] : (T)0;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            int k = k0 + q;
            // |X[k]| > |X[k-1]|, |X[k]| > |X[k+1]|, 10 log10|X[k]| > -30  (:166-168), on squared magnitudes
            bool pk = k >= 1 && k <= M - 2 && pw[q + 1] > pw[q] && pw[q + 1] > pw[q + 2] && pw[q + 1] > (T)1e-6;
            flags |= pk ? (1u << q) : 0u;
        }
    }
    int npk = __popc(flags);
    int incl = npk;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) sm.wsum[warp] = incl;
    __syncthreads();
    int offs = incl - npk;
    for (int i = 0; i < warp; i++) offs += sm.wsum[i];
    if (tid == NT - 1) sm.cnt = offs + npk;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        if (flags & (1u << q)) {
            int k = k0 + q;
            T ssum = 0;                                 // X_fft[k-3:k+3] with python slice semantics (:448)
            if (k >= 3) {
                int hi = k + 3 < M ? k + 3 : M;
                for (int jj = k - 3; jj < hi; jj++) ssum += sm.P[jj];
            }
            T pmv = spl_of<T>(tb.cnorm * ssum);
            T lv = (T)0.367 * (pmv - (T)40 > 0 ? pmv - (T)40 : (T)0);     // :114
            const double zp = zpeakd[k];
            const T zph = (T)zp;
            sm.mz[offs] = zph; sm.mzlo[offs] = (T)(zp - (double)zph); sm.mp[offs] = pmv; sm.ml[offs] = lv;
            offs++;
        }
    }
    __syncthreads();
    const int cnt = sm.cnt;
    T acc[4] = {0, 0, 0, 0};
    for (int m = 0; m < cnt; m++) spread4(acc, zl, zlo, sm.mz[m], sm.mzlo[m], sm.mp[m], sm.ml[m], drop);
#pragma unroll
    for (int j = 0; j < 4; j++) thr[j] = spl_of<T>(acc[j] + tiq[j]);      // :454-456
    __syncthreads();
}


// ---------------------------------------------------------------------------------------------------------------
// fp32 fast threshold curve.  Same quantity as masked_curve (calcBTHR, psychoac.py:431-456) but the fixed-slope parts
// of the spreading function are evaluated by weighted scans over the lines (static weights), the plateau by range
// sums over bins (4-level block sums), and only the upper skirts of maskers louder than 40 dB pairwise -- over the
// lines above them only.  tests/model_analysis.py:curve_v2 is the numpy statement of exactly this procedure.
// Scans: thread t owns lines 4t..4t+3; pairwise part: warp w owns half-chunks w and 2NW-1-w (2 lines per lane each).
// Force-inlined into the single `#pragma unroll 1` loop over the six curves: as a real call its table pointers (kernel
// parameters) were reached through generic loads.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// SPL(intensity) with lg2.approx (abs error ~7e-7 dB)
__device__ __forceinline__ float spl_fast(float inten) {
    inten = fmaxf(inten, 2.511886431509582e-13f);
    return fmaxf(fmaf(lg2_approx(inten), 3.0102999566398120f, 96.0f), -30.0f);
}

__device__ __forceinline__ float spl_any(float i) { return spl_fast(i); }
__device__ __forceinline__ double spl_any(double i) { return spl_of<double>(i); }

template <int LOGM>
__device__ __forceinline__ float4 masked_curve_fast(AnalysisSmem<float, LOGM, true> &sm, CurveScratch<LOGM> &cs, const float2 *F, int tap, float drop,
                                                 const DevTables<float> *tbp, const FastTables *ftp, int hcA, int hcB, const float *linesSrc) {
    constexpr int M = 1 << LOGM, NT = M / 4, NW = NT / 32;
    const DevTables<float> &tb = *tbp;
    const FastTables &ft = *ftp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float K = 0.33219280948873623f;           // log2(10)/10
    auto &fs = sm.fs;
    const int lb0 = 64 * hcA + 2 * lane, lb1 = 64 * hcB + 2 * lane;      // first of this lane's two lines in the warp's two half-chunks
    const int k0 = 4 * tid;
    // 1. power spectrum (optionally of the Hann-tapped spectrum)
    {
        const float2 hw = tb.hann_w, hwc = cconj(tb.hann_w);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int k = tid + NT * j;
            float2 v = tap ? hann_tap<float>(F, k, hw, hwc) : F[k];
            sm.P[k] = v.x * v.x + v.y * v.y;
        }
    }
    const uint2 beu = ft.binEU[tid];                       // static rows, fetched now so that their latency hides behind the barrier
    const float *w = ft.sD + tid, *wa = ft.sA + tid;
    const float wl0 = w[0], wl1 = w[NT], wl2 = w[2 * NT];
    __syncthreads();
    // last curve only: its spectrum (W) is consumed, so the block's scaled L/R lines (k_mdct_enc's output, first read in section F)
    // start travelling into W now, behind the rest of this curve
    if (linesSrc) {
        float *dst = reinterpret_cast<float *>(&sm.W[0][0]);
        for (int e = tid; e < 2 * M / 4; e += NT) cp_async16_a(dst + 4 * e, linesSrc + 4 * e);
        cp_async_commit_a();
    }
    // 2. findpeaks on this thread's 4 bins (psychoac.py:158-191), masker intensities (:448), the ordered list of the loud maskers, and
    //    four scans over the bins through one shuffle ladder: loud count (int), plateau prefix (double), lower skirts (descending,
    //    all maskers), quiet upper skirts (ascending)
    unsigned loudf = 0;
    float c0s[4], ups[4], Aq[4], Qq[4];
    {
        const float4 pa4 = *reinterpret_cast<const float4 *>(&sm.P[k0 >= 4 ? k0 - 4 : 0]);      // P[k0-4 .. k0-1] (k0 = 0: unused)
        const float4 pb4 = *reinterpret_cast<const float4 *>(&sm.P[k0]);
        const float4 pc4 = *reinterpret_cast<const float4 *>(&sm.P[k0 + 4]);                     // P[M .. M+7] are kept zero
        const float pw[10] = {k0 ? pa4.y : 0.f, k0 ? pa4.z : 0.f, k0 ? pa4.w : 0.f, pb4.x, pb4.y, pb4.z, pb4.w, pc4.x, pc4.y, pc4.z};   // P[k0-3 .. k0+6]
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int k = k0 + q;
            const float pc = pw[q + 3];
            const bool pk = k >= 1 && k <= M - 2 && pc > pw[q + 2] && pc > pw[q + 4] && pc > 1e-6f;
            float Av = 0.f, Qv = 0.f;
            c0s[q] = 0.f; ups[q] = 0.f;
            if (pk) {
                // X_fft[k-3:k+3] with python slice semantics (psychoac.py:448): empty for k < 3
                float ssum = k >= 3 ? ((((pw[q] + pw[q + 1]) + pw[q + 2]) + pw[q + 3]) + pw[q + 4]) + pw[q + 5] : 0.f;
                const float pmv = spl_fast(tb.cnorm * ssum);
                const float c0 = (pmv - drop - 96.0f) * K;
                const float lev = 0.367f * fmaxf(pmv - 40.0f, 0.f);
                Av = ex2_approx(c0);
                if (lev > 0.f) {
                    const int eu = (int)((q < 2 ? beu.x : beu.y) >> (16 * (q & 1))) & 0xffff;
                    if (eu < M) { loudf |= 1u << q; ups[q] = (lev - 27.0f) * K; c0s[q] = c0; }
                } else Qv = Av;
            }
            Aq[q] = Av; Qq[q] = Qv;
        }
    }
    int nl = __popc(loudf), incl = nl;
    // plateau intensities: prefix sum in DOUBLE (range sums become differences of prefixes; in fp32 the difference would carry the
    // rounding residue of every louder masker below the window)
    const double A0 = (double)Aq[0], A1 = A0 + (double)Aq[1], A2 = A1 + (double)Aq[2], A3 = A2 + (double)Aq[3];
    double dincl = A3;
    // descending scan over the bins: SD[k] = sum_{k' >= k} A_k' 2^{dn (zp_k' - zp_k)}; ascending: SA[k] = sum_{k' <= k} Q_k' 2^{dn (zp_k - zp_k')}
    const float a3 = Aq[3];
    const float a2 = fmaf(a3, wl2, Aq[2]);
    const float a1 = fmaf(a2, wl1, Aq[1]);
    const float a0 = fmaf(a1, wl0, Aq[0]);
    const float b0 = Qq[0];
    const float b1 = fmaf(b0, wl0, Qq[1]);
    const float b2 = fmaf(b1, wl1, Qq[2]);
    const float b3 = fmaf(b2, wl2, Qq[3]);
    float g = a0, h = b3;
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int o = 1 << s;
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        const double dv = __shfl_up_sync(0xffffffffu, dincl, o);
        const float hv = __shfl_up_sync(0xffffffffu, h, o);
        const float gv = __shfl_down_sync(0xffffffffu, g, o);
        if (lane >= o) { incl += v; dincl += dv; h = fmaf(hv, wa[s * NT], h); }
        if (lane + o < 32) g = fmaf(gv, w[(3 + s) * NT], g);
    }
    float gn = __shfl_down_sync(0xffffffffu, g, 1);
    float hp = __shfl_up_sync(0xffffffffu, h, 1);
    if (lane == 31) { gn = 0.f; sm.wsum[warp] = incl; fs.wtot[warp] = dincl; fs.totA[warp] = h; }
    if (lane == 0) { hp = 0.f; fs.totD[warp] = g; }
    __syncthreads();
    // 3. cross-warp carries; the scans' values, the plateau prefix and the loud list go to shared memory
    {
        int offs = incl - nl;
        double dbase = dincl - A3;
        // (the three carry loops of this step were also tried unrolled with a guard instead of a warp-dependent trip count: 1520 -> 1538 ms)
        for (int i = 0; i < warp; i++) { offs += sm.wsum[i]; dbase += fs.wtot[i]; }
        cs.Sp[k0] = dbase; cs.Sp[k0 + 1] = dbase + A0; cs.Sp[k0 + 2] = dbase + A1; cs.Sp[k0 + 3] = dbase + A2;
        if (tid == NT - 1) cs.Sp[M] = dbase + A3;
        fs.loudBase[tid] = (unsigned short)offs;
        fs.loudFlag[tid] = (unsigned char)loudf;
        if (loudf) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                if (loudf & (1u << q)) {
                    const float2 zk = ft.binZ[k0 + q];
                    const double Bd = ((double)c0s[q] - 0.5 * (double)ups[q]) - (double)ups[q] * ((double)zk.x + (double)zk.y);
                    const float Bh = (float)Bd;
                    const int eu = (int)((q < 2 ? beu.x : beu.y) >> (16 * (q & 1))) & 0xffff;
                    fs.loud[offs] = make_float4(Bh, ups[q], (float)(Bd - (double)Bh), __int_as_float(eu));
                    offs++;
                }
            }
        }
        if (tid == NT - 1) { fs.loudBase[NT] = (unsigned short)offs; fs.loudFlag[NT] = 0; }
        float C = 0.f;
        for (int c2 = NW - 1; c2 > warp; c2--) C = fmaf(C, ft.omD[c2], fs.totD[c2]);
        float inc = fmaf(C, w[8 * NT], gn);
        *reinterpret_cast<float4 *>(&cs.SD[k0]) = make_float4(fmaf(inc, w[9 * NT], a0), fmaf(inc, w[10 * NT], a1), fmaf(inc, w[11 * NT], a2), fmaf(inc, w[12 * NT], a3));
        C = 0.f;
        for (int c2 = 0; c2 < warp; c2++) C = fmaf(C, ft.omA[c2], fs.totA[c2]);
        inc = fmaf(C, wa[5 * NT], hp);
        *reinterpret_cast<float4 *>(&cs.SA[k0]) = make_float4(fmaf(inc, wa[6 * NT], b0), fmaf(inc, wa[7 * NT], b1), fmaf(inc, wa[8 * NT], b2), fmaf(inc, wa[9 * NT], b3));
    }
    // the static records of the first half-chunk's two lines are requested before the barrier: their latency hides behind it
    // (consumed right after it, they cost 1 % of the step)
    const float4 pre0 = ft.lineRec[lb0], pre1 = ft.lineRec[lb0 + 1];
    const float4 prez = *reinterpret_cast<const float4 *>(&ft.lineZ[lb0]);
    __syncthreads();
    // 4. upper skirts of the loud maskers, pairwise, only over lines above them.  Warp w owns the 64-line half-chunks w
    //    and 2*NW-1-w, two adjacent lines per lane in each, so that every warp sees the same number of (masker, line)
    //    pairs when maskers are spread evenly over the bins.
    //    Culling (warp-uniform): an upper skirt decays along the lines, so its largest value inside a half-chunk is at most
    //    its value extrapolated to the half-chunk's first line.  Maskers whose bound lies 30 bits below the smallest partial
    //    threshold of the half-chunk (lower skirts + plateaus + quiet skirts + threshold in quiet are already in) cannot change
    //    a float sum: each lane tests one masker, a ballot keeps the survivors (about a third on the synthetic corpus), and the
    //    pairwise loop walks only those.  <= 512 maskers x 2^-26: 8e-6 relative, 3e-5 dB (the 1e-5 budget on SMR is 1e-4 dB).
    float acc[4];
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
        const int l0 = hh ? lb1 : lb0;                    // first of this lane's two lines in half-chunk hh
        // partial thresholds of the lane's two lines: lower skirts + quiet upper skirts from one scan entry each and a static factor,
        // plateau as a range sum over the bins within half a Bark, threshold in quiet (psychoac.py:437,452-454)
        float x0, x1;
        {
            // second half-chunk: this thread's own copy in shared memory (staged once per CTA; the global loads, consumed at once, cost 1 %)
            const float4 r0 = hh ? fs.stash[0][tid] : pre0, r1 = hh ? fs.stash[1][tid] : pre1;
            const unsigned w0 = __float_as_uint(r0.x), w1 = __float_as_uint(r1.x);
            const int pa0 = (int)(w0 & 0xffffu), pb0 = (int)(w0 >> 16), pa1 = (int)(w1 & 0xffffu), pb1 = (int)(w1 >> 16);
            const float pl0 = (float)(cs.Sp[pb0] - cs.Sp[pa0]), pl1 = (float)(cs.Sp[pb1] - cs.Sp[pa1]);
            x0 = fmaf(cs.SD[min(pb0, M - 1)], r0.y, fmaf(cs.SA[max(pa0 - 1, 0)], r0.z, pl0)) + r0.w;
            x1 = fmaf(cs.SD[min(pb1, M - 1)], r1.y, fmaf(cs.SA[max(pa1 - 1, 0)], r1.z, pl1)) + r1.w;
        }
        const float4 zz = hh ? *reinterpret_cast<const float4 *>(&ft.lineZ[l0]) : prez;      // l0 is even: (z0_hi, z0_lo, z1_hi, z1_lo)
        const float z0 = zz.x, z0l = zz.y;
        const float dz01 = (zz.z - zz.x) + (zz.w - zz.y);          // Bark gap to the lane's second line
        const uint32_t kU = ft.kUhc[hh ? hcB : hcA];      // bins whose upper skirt starts at or before the half-chunk's first / last line (static)
        // loud maskers below a bin k = base of its group of four + the loud ones among the group's bins below k
        auto loudBelow = [&](unsigned k) { return (int)fs.loudBase[k >> 2] + __popc((unsigned)fs.loudFlag[k >> 2] & ((1u << (k & 3u)) - 1u)); };
        const int mfull = loudBelow(kU & 0xffffu);            // upper skirt starts at or before the half-chunk
        const int mhi = loudBelow(kU >> 16);                  // ... at or before its last line
        if (mhi > 0) {
            // positive floats order like their bit patterns: one REDUX gives the smallest partial threshold of the half-chunk
            const float lmin = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(fminf(x0, x1))));
            const float cut = lg2_approx(lmin) - 26.01f;
            const float zc = __shfl_sync(0xffffffffu, z0, 0);             // Bark position of the half-chunk's first line
            for (int base = 0; base < mhi; base += 32) {
                const int mm = base + lane;
                bool keep = false;
                if (mm < mhi) {
                    const float2 pc = *reinterpret_cast<const float2 *>(&fs.loud[mm]);       // (B_hi, up)
                    keep = fmaf(pc.y, zc, pc.x) >= cut;
                }
                const unsigned mask = __ballot_sync(0xffffffffu, keep);
                const int nf = mfull - base;
                const unsigned fullMask = nf >= 32 ? 0xffffffffu : (nf > 0 ? (1u << nf) - 1u : 0u);
                unsigned mk = mask & fullMask;
                // four survivors per trip (independent loads and exponentials); a trip's spare slots read the sentinel, which adds 0
                constexpr int SENT = M / 2;
                while (mk) {
                    const int j0 = __ffs(mk) - 1; mk &= mk - 1;
                    const int j1 = __ffs(mk) - 1; mk &= mk - 1;          // __ffs(0) - 1 = -1 once the mask is empty
                    const int j2 = __ffs(mk) - 1; mk &= mk - 1;
                    const int j3 = __ffs(mk) - 1; mk &= mk - 1;
                    const float4 p0 = fs.loud[base + j0];
                    const float4 p1 = fs.loud[j1 < 0 ? SENT : base + j1];
                    const float4 p2 = fs.loud[j2 < 0 ? SENT : base + j2];
                    const float4 p3 = fs.loud[j3 < 0 ? SENT : base + j3];
                    const float e00 = fmaf(p0.y, z0, p0.x) + fmaf(p0.y, z0l, p0.z), e01 = fmaf(p0.y, dz01, e00);
                    const float e10 = fmaf(p1.y, z0, p1.x) + fmaf(p1.y, z0l, p1.z), e11 = fmaf(p1.y, dz01, e10);
                    const float e20 = fmaf(p2.y, z0, p2.x) + fmaf(p2.y, z0l, p2.z), e21 = fmaf(p2.y, dz01, e20);
                    const float e30 = fmaf(p3.y, z0, p3.x) + fmaf(p3.y, z0l, p3.z), e31 = fmaf(p3.y, dz01, e30);
                    const float t00 = ex2_approx(e00), t01 = ex2_approx(e01), t10 = ex2_approx(e10), t11 = ex2_approx(e11);
                    const float t20 = ex2_approx(e20), t21 = ex2_approx(e21), t30 = ex2_approx(e30), t31 = ex2_approx(e31);
                    x0 = ((x0 + t00) + t10) + t20 + t30;                   // ascending masker order, as the one-by-one loop
                    x1 = ((x1 + t01) + t11) + t21 + t31;
                }
                // maskers whose upper skirt starts INSIDE the half-chunk: per line a test against the skirt's first line.  Two per trip
                // (the second slot reads the sentinel when the mask runs out: exponent -inf, first line M): their loads overlap
                mk = mask & ~fullMask;
                while (mk) {
                    const int ja = __ffs(mk) - 1; mk &= mk - 1;
                    const int jb = __ffs(mk) - 1; mk &= mk - 1;
                    const float4 pa = fs.loud[base + ja];
                    const float4 pb = fs.loud[jb < 0 ? SENT : base + jb];
                    const int eua = __float_as_int(pa.w), eub = __float_as_int(pb.w);
                    const float ea0 = fmaf(pa.y, z0, pa.x) + fmaf(pa.y, z0l, pa.z), ea1 = fmaf(pa.y, dz01, ea0);
                    const float eb0 = fmaf(pb.y, z0, pb.x) + fmaf(pb.y, z0l, pb.z), eb1 = fmaf(pb.y, dz01, eb0);
                    const float ta0 = ex2_approx(ea0), ta1 = ex2_approx(ea1), tb0 = ex2_approx(eb0), tb1 = ex2_approx(eb1);
                    x0 += (l0 >= eua) ? ta0 : 0.f;
                    x1 += (l0 + 1 >= eua) ? ta1 : 0.f;
                    x0 += (l0 >= eub) ? tb0 : 0.f;                         // ascending masker order, as before
                    x1 += (l0 + 1 >= eub) ? tb1 : 0.f;
                }
            }
        }
        acc[2 * hh] = spl_fast(x0);
        acc[2 * hh + 1] = spl_fast(x1);
    }
    // no barrier here: the next curve's step 1 only writes P (last read in step 2) and every later step sits behind its own barrier;
    // the caller fences before anything rewrites the scratch buffer or the loud list out of turn
    return make_float4(acc[0], acc[1], acc[2], acc[3]);
}

// MDCT_ONLY = true is the stage instantiation behind pac_mdct_batch: sections A + C (PCM -> fractions, SineWindow, MDCT,
// overall scale), writing the scaled L/R lines and the overall scales, nothing else -- the window+MDCT stage of the encoder
// measured by itself (SURVEY.md 8d asks for the MDCT and SMR roofline fractions separately).
template <typename T, int LOGM, bool MDCT_ONLY = false>
__global__ void __launch_bounds__((1 << LOGM) / 4, sizeof(T) == 4 ? PAC_ANALYSIS_CTAS : 1)
k_analysis(const __grid_constant__ AnalysisArgs<T> a) {
    using S = AnalysisSmem<T, LOGM, sizeof(T) == 4>;
    using T2 = typename Vec2<T>::type;
    constexpr int M = S::M, N = S::N, NT = S::NT, H = M / 2, NW = NT / 32;
    // fp32: four CTAs per SM need 4 x (dynamic + 1 KB reserved) <= 228 KB
    static_assert(sizeof(T) == 8 || sizeof(S) <= 57344, "fp32 analysis must fit four CTAs per SM");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S &sm = *reinterpret_cast<S *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevTables<T> &tb = a.tab;
    const int NB = a.bands.nBands;
    constexpr bool FAST = sizeof(T) == 4;
    constexpr bool R16 = FAST && LOGM == 10;          // the fp32 batch of four 1024-point FFTs runs register-blocked (fft.cuh)
    constexpr int ROW = S::ROW;
    T *xt = reinterpret_cast<T *>(&sm.XF[0][0]);      // x[ch][n] at xt[ch*XS + XI(n)]
    constexpr int XS = 2 * ROW;
    // fp32: sample pair m = (x[2m], x[2m+1]) is complex element m of the in-place raw FFT, stored at its padded position
    auto XI = [](int n) { return FAST ? n + 2 * (n >> 5) : n; };
    auto PI = [](int m) { return fft_pad<FAST>(m); };

    // per-thread constants for its 4 lines i = tid + NT*j
    // line ownership: fp64 (direct evaluation) thread t owns lines t + NT*j; fp32 (scan-based evaluation): warp w owns the
    // 64-line half-chunks w and 2*NW-1-w, two adjacent lines per lane in each (see masked_curve_fast step 4).
    // With the culling of step 4 the surviving (masker, half-chunk) pairs per half-chunk no longer form a triangle: measured on the
    // synthetic corpus [4.7 5.6 5.9 6.8 7.2 7.3 7.8 8.1 8.2 8.8 9.7 10.4 9.8 5.8 0.5 0.0] maskers for half-chunks 0..15 (the top ones sit
    // under the steep threshold in quiet), so heavy half-chunks are paired with light ones (max 14.6 per warp instead of 17.6).
    const int hcA = NW == 8 ? (warp < 2 ? 11 + warp : 12 - warp) : warp;                     // 11 12 10 9 8 7 6 5
    const int hcB = NW == 8 ? (warp < 2 ? 15 - warp : (warp == 4 ? 13 : (warp < 4 ? warp - 2 : warp - 3))) : 2 * NW - 1 - warp;   // 15 14 0 1 13 2 3 4
    const int lineBase[2] = {64 * hcA + 2 * lane, 64 * hcB + 2 * lane};
    auto LI = [&](int j) { return FAST ? lineBase[j >> 1] + (j & 1) : tid + NT * j; };
    T zl[4], zlo[4], tiq[4], mld[4];
    int bnd[4];
    if constexpr (!FAST) {          // (fp32: section F fetches its two per-line constants when it needs them -- eight registers less to
#pragma unroll                      //  carry through the six curves of a 64-register kernel)
        for (int j = 0; j < 4; j++) {
            int i = LI(j);
            const double zd = a.tabd.zline[i];
            zl[j] = (T)zd; zlo[j] = (T)(zd - (double)zl[j]);
            tiq[j] = tb.tiq[i]; mld[j] = tb.mld[i]; bnd[j] = tb.band_of_line[i];
        }
    }

    if (tid < 8) sm.P[M + tid] = 0;
    if constexpr (FAST && !MDCT_ONLY) { if (tid == 8) sm.fs.loud[M / 2] = make_float4(-INFINITY, 0.f, 0.f, __int_as_float(M)); }
    if constexpr (FAST && !MDCT_ONLY) { sm.fs.stash[0][tid] = a.ft.lineRec[lineBase[1]]; sm.fs.stash[1][tid] = a.ft.lineRec[lineBase[1] + 1]; }
    for (int64_t w = blockIdx.x; w < a.nwork; w += gridDim.x) {
        // 32-bit division when the tile allows it (a 64-bit division is a ~100-instruction subroutine, paid by every CTA)
        const int s = a.nwork <= 0xffffffffll ? (int)((uint32_t)w / (uint32_t)a.nb) : (int)(w / a.nb);
        const int b = a.b0 + (int)(w - (int64_t)s * a.nb);
        if (a.poisonOn) {             // results must not depend on what a previous block left in shared memory
            __syncthreads();
            for (uint32_t i = tid; i < a.smemWords; i += NT) reinterpret_cast<uint32_t *>(smem_raw)[i] = a.poison;
            __syncthreads();
            if (tid < 8) sm.P[M + tid] = 0;
            if constexpr (FAST && !MDCT_ONLY) { if (tid == 8) sm.fs.loud[M / 2] = make_float4(-INFINITY, 0.f, 0.f, __int_as_float(M)); }
            if constexpr (FAST && !MDCT_ONLY) { sm.fs.stash[0][tid] = a.ft.lineRec[lineBase[1]]; sm.fs.stash[1][tid] = a.ft.lineRec[lineBase[1] + 1]; }
            __syncthreads();
        }
        // ------------------------------------------------ A. load one 2048-sample stereo window
        if (a.pcm) {
            const int64_t ns = a.nSamples[s];
            const int64_t nblk = (ns + M - 1) / M + 1;
            if (b >= nblk) continue;                     // uniform for the CTA
            const int *p32 = reinterpret_cast<const int *>(a.pcm) + (int64_t)s * a.strideSamples;
            const int64_t base = (int64_t)(b - 1) * M;
#pragma unroll
            for (int j = 0; j < N / NT; j++) {
                int n = tid + NT * j;
                int64_t si = base + n;
                int v = (si >= 0 && si < ns) ? __ldg(p32 + si) : 0;
#pragma unroll
                for (int ch = 0; ch < 2; ch++) {
                    int c = ch ? (v >> 16) : (int)(short)(v & 0xffff);
                    int code = c < 0 ? -c : c;
                    if (code & 32768) code -= 32768;          // -32768 dequantises to 0 (quantize.py:133-138)
                    T f;                                      // quantize.py:141: 2 * code / 65535
                    if constexpr (FAST) f = (float)(2 * code) * 1.5259021896696422e-05f;      // (psychoacoustic input only in fp32 mode: one rounding more, no division)
                    else f = (T)(2 * code) / (T)65535;
                    xt[ch * XS + XI(n)] = c < 0 ? -f : f;
                }
            }
        } else {
            const double *src = a.blocks + w * 2 * N;
#pragma unroll
            for (int j = 0; j < 2 * N / NT; j++) {
                int e = tid + NT * j;
                int ch = e / N, n = e - ch * N;
                xt[ch * XS + XI(n)] = (T)src[e];
            }
        }
        __syncthreads();
        int osc0 = 0, osc1 = 0;
        auto sectionC = [&]() {
            // ------------------------------------------------ C. SineWindow (in place) + MDCT + overall scale
            const bool inPlace = tb.winInPlace != 0;              // SineWindow multiplies the block itself (window.py:37); KBDWindow a copy (:64)
            if constexpr (!FAST) {
                if (inPlace) {
                    for (int e = tid; e < 2 * N; e += NT) {
                        int ch = e / N, n = e - ch * N;
                        xt[ch * XS + n] *= tb.sinw[n];
                    }
                    __syncthreads();
                }
            }
            T mx[2] = {0, 0};
            {
                for (int e = tid; e < 2 * H; e += NT) {            // fold to M/2 complex points per channel
                    int ch = e / H, n = e - ch * H;
                    const T *xc = xt + ch * XS;
                    auto x = [&](int i) -> T { return inPlace ? xc[i] : xc[i] * tb.sinw[i]; };
                    int m0 = 2 * n, m1 = M - 1 - 2 * n;
                    T u0 = m0 < H ? -x(3 * H - 1 - m0) - x(3 * H + m0) : x(m0 - H) - x(2 * H - 1 - (m0 - H));
                    T u1 = m1 < H ? -x(3 * H - 1 - m1) - x(3 * H + m1) : x(m1 - H) - x(2 * H - 1 - (m1 - H));
                    sm.W[ch][n] = cmul(mk2<T>(u0, u1), tb.mdct_pre[n]);
                }
                __syncthreads();
                fft_dif<T, LOGM - 1, NT>(&sm.W[0][0], 2, ROW, tb.tw, 2);
                for (int e = tid; e < 2 * H; e += NT) {
                    int ch = e / H, k = e - ch * H;
                    T2 y = cmul(sm.W[ch][fft_pos<LOGM - 1>(k)], tb.mdct_post[k]);
                    T v0 = ((T)2 / (T)N) * y.x, v1 = -((T)2 / (T)N) * y.y;
                    sm.Lb[ch][2 * k] = v0;
                    sm.Lb[ch][M - 1 - 2 * k] = v1;
                    T m = fmax(fabs(v0), fabs(v1));
                    if (ch == 0) mx[0] = fmax(mx[0], m); else mx[1] = fmax(mx[1], m);
                }
            }
            mx[0] = warp_max(mx[0]); mx[1] = warp_max(mx[1]);
            if (lane == 0) { sm.red[warp] = mx[0]; sm.red[32 + warp] = mx[1]; }
            __syncthreads();
            if (tid < 2) {
                T m = 0;
                for (int i = 0; i < NW; i++) m = fmax(m, sm.red[32 * tid + i]);
                sm.oscale[tid] = scale_factor((double)m, a.nScaleBits, 5);   // codec.py:245: ScaleFactor(maxLine, nScaleBits) with the default nMantBits=5
            }
            __syncthreads();
            osc0 = sm.oscale[0]; osc1 = sm.oscale[1];
            for (int e = tid; e < 2 * M; e += NT) {
                int ch = e / M, i = e - ch * M;
                sm.Lb[ch][i] *= (T)(1 << (ch ? osc1 : osc0));      // codec.py:246
            }
        };
        if constexpr (FAST) {
            // fp32 mode: window + MDCT + overall scale ran as their own fp64 kernel (mdct.cuh: k_mdct) and left the scaled L/R lines in
            // a.lines and the scales in a.oscale.  The lines travel into shared memory by cp.async during the last curve
            // (masked_curve_fast; section F is their first reader, and overwrites a.lines in place with the LRMS-selected lines).
            osc0 = a.oscale[w * 2]; osc1 = a.oscale[w * 2 + 1];
        }
        if constexpr (MDCT_ONLY) sectionC();
        if constexpr (MDCT_ONLY) {
            __syncthreads();
            for (int e = tid; e < 2 * M; e += NT) a.lines[w * 2 * M + e] = sm.Lb[e / M][e % M];
            if (tid < 2) a.oscale[w * 2 + tid] = (uint8_t)(tid ? osc1 : osc0);
            __syncthreads();
            continue;
        }
        // ------------------------------------------------ B. raw FFTs -> LRMS decision (codec.py:96-102)
        // fp64: raw FFTs of L, R in W.  fp32: the raw FFTs run IN PLACE in XF and, in the same batch of four, the FFTs of the
        // sine*Hann windowed channels in W (the fp32 MDCT has already consumed x).
        T2(*RAW)[ROW] = FAST ? sm.XF : sm.W;
        for (int e = tid; e < 2 * M; e += NT) {
            int ch = e / M, m = e - ch * M;
            const T2 x2 = sm.XF[ch][PI(m)];               // (x[2m], x[2m+1])
            if constexpr (FAST) {
                const T2 s2 = reinterpret_cast<const T2 *>(tb.sinw)[m], h2 = reinterpret_cast<const T2 *>(tb.hann)[m];
                if (tb.winInPlace) sm.W[ch][PI(m)] = mk2<T>((x2.x * s2.x) * h2.x, (x2.y * s2.y) * h2.y);      // window.py:37 then psychoac.py:428
                else sm.W[ch][PI(m)] = mk2<T>(x2.x * h2.x, x2.y * h2.y);                                     // KBDWindow left the block alone
            } else {
                sm.W[ch][m] = x2;
            }
        }
        if (tid == 0) sm.lrms = 0;
        __syncthreads();
        if constexpr (R16) fft1024_r16<NT, true>(&sm.XF[0][0], 4, ROW, tb.tw);      // register-blocked radix 16 x 16 x 4
        else if constexpr (FAST) fft_dif<T, LOGM, NT, true>(&sm.XF[0][0], 4, ROW, tb.tw, 1);
        else fft_dif<T, LOGM, NT>(&sm.W[0][0], 2, ROW, tb.tw, 1);
        for (int bd = warp; bd < NB; bd += NW) {
            T dr = 0, di = 0, sr = 0, si = 0;
            for (int k = a.bands.lo[bd] + lane; k < a.bands.lo[bd + 1]; k += 32) {
                T2 l = rfft_split<T, LOGM, FAST, R16>(RAW[0], k, tb.tw_split);
                T2 r = rfft_split<T, LOGM, FAST, R16>(RAW[1], k, tb.tw_split);
                T l2r = l.x * l.x - l.y * l.y, l2i = l.x * l.y + l.y * l.x;
                T r2r = r.x * r.x - r.y * r.y, r2i = r.x * r.y + r.y * r.x;
                dr += l2r - r2r; di += l2i - r2i;
                sr += l2r + r2r; si += l2i + r2i;
            }
            dr = warp_sum(dr); di = warp_sum(di); sr = warp_sum(sr); si = warp_sum(si);
            if (lane == 0) {
                if constexpr (FAST) {
                    if (dr * dr + di * di < 0.64f * (sr * sr + si * si)) atomicOr(&sm.lrms, 1u << bd);
                } else {
                    double dd = hypot((double)dr, (double)di), ss = hypot((double)sr, (double)si);
                    if (dd < 0.8 * ss) atomicOr(&sm.lrms, 1u << bd);
                }
            }
        }
        __syncthreads();
        if constexpr (FAST) {
            for (int e = tid; e < 2 * (M + 1); e += NT) {         // F1[ch][k], k = 0..M, into XF (the raw spectra are dead)
                int ch = e / (M + 1), k = e - ch * (M + 1);
                sm.XF[ch][k] = rfft_split<T, LOGM, true, R16>(sm.W[ch], k, tb.tw_split);
            }
            __syncthreads();
        }
        if constexpr (!FAST && !MDCT_ONLY) sectionC();
        if constexpr (!FAST) {
            // ------------------------------------------------ D. Hann on the sine-windowed data (psychoac.py:428) + FFT
            for (int e = tid; e < 2 * M; e += NT) {
                int ch = e / M, m = e - ch * M;
                T2 x2 = sm.XF[ch][m];
                const T2 h2 = reinterpret_cast<const T2 *>(tb.hann)[m];
                sm.W[ch][m] = mk2<T>(x2.x * h2.x, x2.y * h2.y);
            }
            __syncthreads();
            fft_dif<T, LOGM, NT>(&sm.W[0][0], 2, ROW, tb.tw, 1);
            for (int e = tid; e < 2 * (M + 1); e += NT) {         // F1[ch][k], k = 0..M
                int ch = e / (M + 1), k = e - ch * (M + 1);
                sm.XF[ch][k] = rfft_split<T, LOGM>(sm.W[ch], k, tb.tw_split);
            }
            __syncthreads();
        } else {
            __syncthreads();
        }
        // F2_M, F2_S = Hann-taps of (F1_L +- F1_R)/2   (psychoac.py:549 then :428 again)
        const T2 hw = tb.hann_w, hwc = cconj(tb.hann_w);
        auto computeF2 = [&]() {
            for (int e = tid; e < 2 * (M + 1); e += NT) {
                int c = e / (M + 1), k = e - c * (M + 1);
                T sg = c ? (T)-1 : (T)1;
                int km = k == 0 ? 1 : k - 1, kp = k == M ? M - 1 : k + 1;
                T2 f0 = sm.XF[0][k], f1 = sm.XF[1][k];
                T2 g0 = sm.XF[0][km], g1 = sm.XF[1][km];
                T2 h0 = sm.XF[0][kp], h1 = sm.XF[1][kp];
                T2 fc = mk2<T>((f0.x + sg * f1.x) / 2, (f0.y + sg * f1.y) / 2);
                T2 fm = mk2<T>((g0.x + sg * g1.x) / 2, (g0.y + sg * g1.y) / 2);
                T2 fp = mk2<T>((h0.x + sg * h1.x) / 2, (h0.y + sg * h1.y) / 2);
                if (k == 0) fm = cconj(fm);
                if (k == M) fp = cconj(fp);
                T2 t = cadd(cmul(hw, fm), cmul(hwc, fp));
                sm.W[c][k] = mk2<T>((T)0.5 * fc.x - (T)0.25 * t.x, (T)0.5 * fc.y - (T)0.25 * t.y);
            }
            __syncthreads();
        };
        if constexpr (!FAST) computeF2();
        // ------------------------------------------------ E. six masked-threshold curves
        T thr[6][4];
        if constexpr (FAST) {
            // The dense per-bin scratch of a curve lives in the spectrum buffer that is dead at that point: the L and R curves read
            // F1 from XF while W (the consumed FFT work area) is free; then F2 is formed in W and XF becomes the scratch.
            static_assert(sizeof(CurveScratch<LOGM>) <= sizeof(sm.W) && sizeof(CurveScratch<LOGM>) <= sizeof(sm.XF), "scratch must fit a spectrum buffer");
#pragma unroll 1
            for (int c = 0; c < 6; c++) {
                if (c == 2) { __syncthreads(); computeF2(); }      // W (scratch of the L/R curves, read until the end of their step 4) becomes F2
                CurveScratch<LOGM> &cs = *reinterpret_cast<CurveScratch<LOGM> *>(c < 2 ? &sm.W[0][0] : &sm.XF[0][0]);
                const float2 *src = c < 2 ? sm.XF[c] : sm.W[c & 1];
                const float4 r = masked_curve_fast<LOGM>(sm, cs, src, c >= 4, c < 4 ? 15.f : 0.f, &a.tab, &a.ft, hcA, hcB,
                                                         c == 5 ? reinterpret_cast<const float *>(a.lines) + w * 2 * M : nullptr);
                thr[c][0] = r.x; thr[c][1] = r.y; thr[c][2] = r.z; thr[c][3] = r.w;
            }
        } else {
            masked_curve<T, LOGM, S>(sm, tb, [&](int k) { return sm.XF[0][k]; }, (T)15, zl, zlo, tiq, a.tabd.zpeak, thr[0]);   // BTHR_L  :540
            masked_curve<T, LOGM, S>(sm, tb, [&](int k) { return sm.XF[1][k]; }, (T)15, zl, zlo, tiq, a.tabd.zpeak, thr[1]);   // BTHR_R  :541
            masked_curve<T, LOGM, S>(sm, tb, [&](int k) { return sm.W[0][k]; }, (T)15, zl, zlo, tiq, a.tabd.zpeak, thr[2]);    // BTHR_M  :559
            masked_curve<T, LOGM, S>(sm, tb, [&](int k) { return sm.W[1][k]; }, (T)15, zl, zlo, tiq, a.tabd.zpeak, thr[3]);    // BTHR_S  :560
            masked_curve<T, LOGM, S>(sm, tb, [&](int k) { return hann_tap<T>(sm.W[0], k, hw, hwc); }, (T)0, zl, zlo, tiq, a.tabd.zpeak, thr[4]);   // :561
            masked_curve<T, LOGM, S>(sm, tb, [&](int k) { return hann_tap<T>(sm.W[1], k, hw, hwc); }, (T)0, zl, zlo, tiq, a.tabd.zpeak, thr[5]);   // :562
        }
        // ------------------------------------------------ F. SMR candidates, band maxima, select
        if constexpr (FAST) { cp_async_wait_a(); __syncthreads(); }       // the L/R lines have landed (in W); the last curve's scratch (XF) is dead
        const uint32_t lrms = sm.lrms;
        // V[q][i], q = 0..3 (L,R,M,S), stride M: fp64 in W; fp32 in XF, because W holds the lines there
        T *V = reinterpret_cast<T *>(FAST ? &sm.XF[0][0] : &sm.W[0][0]);
        T(*LB)[M] = nullptr;
        if constexpr (FAST) LB = reinterpret_cast<T(*)[M]>(&sm.W[0][0]); else LB = sm.Lb;
        T outl[2][4];
        if constexpr (FAST) {
#pragma unroll
            for (int j = 0; j < 4; j++) { mld[j] = tb.mld[LI(j)]; bnd[j] = tb.band_of_line[LI(j)]; }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int i = LI(j);
            T xl = LB[0][i], xr = LB[1][i];
            T xm = (xl + xr) / 2, xs = (xl - xr) / 2;                                     // psychoac.py:551
            T sl = spl_any((T)4 * (xl * xl)) - (T)6.02 * (T)osc0;                        // :534
            T sr = spl_any((T)4 * (xr * xr)) - (T)6.02 * (T)osc1;                        // :535
            T sM = spl_any((T)4 * (xm * xm)) - (T)6.02 * (T)osc0;                        // :554
            T sS = spl_any((T)4 * (xs * xs)) - (T)6.02 * (T)osc1;                        // :555
            T mldM = thr[4][j] * mld[j], mldS = thr[5][j] * mld[j];                      // :582-583
            T thrM = fmax(thr[2][j], fmin(thr[3][j], mldS));                             // :591
            T thrS = fmax(thr[3][j], fmin(thr[2][j], mldM));
            V[0 * M + i] = sl - thr[0][j];
            V[1 * M + i] = sr - thr[1][j];
            V[2 * M + i] = sM - thrM;
            V[3 * M + i] = sS - thrS;
            bool ms = (lrms >> bnd[j]) & 1u;
            outl[0][j] = ms ? xm : xl;
            outl[1][j] = ms ? xs : xr;
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int i = LI(j);
            a.lines[(w * 2 + 0) * M + i] = outl[0][j];
            a.lines[(w * 2 + 1) * M + i] = outl[1][j];
        }
        if (a.dbg_mdct) {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int i = LI(j);
                a.dbg_mdct[(w * 2 + 0) * M + i] = LB[0][i];
                a.dbg_mdct[(w * 2 + 1) * M + i] = LB[1][i];
            }
        }
        if (a.dbg_bthr) {
#pragma unroll
            for (int c = 0; c < 6; c++)
#pragma unroll
                for (int j = 0; j < 4; j++) a.dbg_bthr[(w * 6 + c) * M + LI(j)] = thr[c][j];
        }
        __syncthreads();
        for (int bd = warp; bd < NB; bd += NW) {
            const bool ms = (lrms >> bd) & 1u;
            T v0 = -INFINITY, v1 = -INFINITY, m0 = 0, m1 = 0;
            const int lo = a.bands.lo[bd], hi = a.bands.lo[bd + 1];
            for (int i = lo + lane; i < hi; i += 32) {
                T xl = LB[0][i], xr = LB[1][i];
                T c0 = ms ? (xl + xr) / 2 : xl, c1 = ms ? (xl - xr) / 2 : xr;
                v0 = fmax(v0, V[(ms ? 2 : 0) * M + i]);
                v1 = fmax(v1, V[(ms ? 3 : 1) * M + i]);
                m0 = fmax(m0, fabs(c0)); m1 = fmax(m1, fabs(c1));
            }
            v0 = warp_max(v0); v1 = warp_max(v1); m0 = warp_max(m0); m1 = warp_max(m1);
            if (lane == 0) {
                if (hi == lo) { v0 = (T)-96; v1 = (T)-96; }                              // psychoac.py:496-498
                a.smr[(w * 2 + 0) * kMaxBands + bd] = v0;
                a.smr[(w * 2 + 1) * kMaxBands + bd] = v1;
                a.bmax[(w * 2 + 0) * kMaxBands + bd] = m0;
                a.bmax[(w * 2 + 1) * kMaxBands + bd] = m1;
            }
        }
        if (tid == 0) {
            a.lrms[w] = lrms;
            a.oscale[w * 2 + 0] = (uint8_t)osc0;
            a.oscale[w * 2 + 1] = (uint8_t)osc1;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- mono CalcSMRs (psychoac.py:215-318)
template <typename T>
struct SmrMonoArgs {
    const double *data;      // [n][N] time samples (Hann applied here, psychoac.py:225)
    const double *mdct;      // [n][M] lines scaled by 2^scale
    int n, scale, noDrop;
    double *smr;             // [n][kMaxBands] (may be NULL)
    double *thr;             // [n][M] optional masked threshold (getMaskedThreshold)
    DevTables<T> tab;
    DevTables<double> tabd;
    BandInfo bands;
};

template <typename T, int LOGM>
__global__ void __launch_bounds__((1 << LOGM) / 4)
k_calc_smrs(const SmrMonoArgs<T> a) {
    using S = AnalysisSmem<T, LOGM>;
    using T2 = typename Vec2<T>::type;
    constexpr int M = S::M, N = S::N, NT = S::NT, NW = NT / 32 > 0 ? NT / 32 : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    S &sm = *reinterpret_cast<S *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const DevTables<T> &tb = a.tab;
    T zl[4], zlo[4], tiq[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const double zd = a.tabd.zline[tid + NT * j];
        zl[j] = (T)zd; zlo[j] = (T)(zd - (double)zl[j]);
        tiq[j] = tb.tiq[tid + NT * j];
    }
    const int w = blockIdx.x;
    for (int m = tid; m < M; m += NT) {
        T x0 = (T)a.data[(int64_t)w * N + 2 * m] * tb.hann[2 * m];
        T x1 = (T)a.data[(int64_t)w * N + 2 * m + 1] * tb.hann[2 * m + 1];
        sm.W[0][m] = mk2<T>(x0, x1);
    }
    __syncthreads();
    fft_dif<T, LOGM, NT>(&sm.W[0][0], 1, M + 2, tb.tw, 1);
    for (int k = tid; k <= M; k += NT) sm.XF[0][k] = rfft_split<T, LOGM>(sm.W[0], k, tb.tw_split);
    __syncthreads();
    T thr[4];
    masked_curve<T, LOGM, S>(sm, tb, [&](int k) { return sm.XF[0][k]; }, a.noDrop ? (T)0 : (T)15, zl, zlo, tiq, a.tabd.zpeak, thr);
    T *V = reinterpret_cast<T *>(&sm.W[0][0]);
    const T sc = (T)exp2((double)a.scale);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int i = tid + NT * j;
        if (a.thr) a.thr[(int64_t)w * M + i] = (double)thr[j];
        if (a.smr) {
            T tr = (T)a.mdct[(int64_t)w * M + i] / sc;             // :285
            V[i] = spl_of<T>((T)4 * (tr * tr)) - thr[j];            // :286-287, :316
        }
    }
    if (!a.smr) return;
    __syncthreads();
    for (int bd = warp; bd < a.bands.nBands; bd += NW) {
        T v = -INFINITY;
        const int lo = a.bands.lo[bd], hi = a.bands.lo[bd + 1];
        for (int i = lo + lane; i < hi; i += 32) v = fmax(v, V[i]);
        v = warp_max(v);
        if (lane == 0) a.smr[(int64_t)w * kMaxBands + bd] = hi > lo ? (double)v : 0.0;   // :309,314
    }
}

}  // namespace pac
