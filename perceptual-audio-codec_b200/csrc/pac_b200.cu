// pac_b200.cu -- C ABI (include/pac_b200.h) over the sm_100a kernels.  Host side: context, constant tables,
// workspace, tiling of (streams x blocks), H2D/D2H staging.  No CPU fallback anywhere: every entry point
// either launches kernels or fails with an error code.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "analysis.cuh"
#include "common.cuh"
#include "decode.cuh"
#include "mdct.cuh"
#include "pack.cuh"
#include "scan.cuh"
#include "stages.cuh"

using namespace pac;

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static thread_local std::string g_create_error;

// ------------------------------------------------------------------ device buffer (grow-only)
struct DBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename U> U *as() { return reinterpret_cast<U *>(p); }
};

// pinned host buffer (small asynchronous read-backs that must not block the enqueueing thread)
struct HBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 8 + 256;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename U> U *as() { return reinterpret_cast<U *>(p); }
};

template <typename T>
struct TableSet {
    DevTables<T> dev;
    void *mem = nullptr;
};

struct PacCtx {
    int device = 0, precision = 0;
    PacParams p{};
    int M = 0, N = 0, LOGM = 0;
    BandInfo bands{};
    int32_t nLines[kMaxBands]{};
    EncConsts ec{};
    cudaStream_t stream = nullptr;
    cudaStream_t ownStream = nullptr;
    cudaStream_t sA = nullptr, sB = nullptr;      // internal streams: analysis / scan+pack of consecutive tiles overlap
    cudaStream_t sC = nullptr;                    // copy stream: H2D of the next stream group / D2H of the previous one
    cudaStream_t sD = nullptr;                    // k_drain: a tile's finished bytes -> the caller's pinned host image
    cudaStream_t sM = nullptr;                    // fp32 mode: window+MDCT of the tiles ahead (k_mdct_enc), beside the analysis of the current one
    cudaEvent_t evH[2] = {nullptr, nullptr}, evD[2] = {nullptr, nullptr};
    DBuf w_pcm2, w_out2;
    cudaEvent_t evStart = nullptr;
    std::vector<cudaEvent_t> evA, evB, evG, evS[2], evM;
    HBuf hostState;
    cudaStream_t launchStream = nullptr;          // stream the launch helpers currently target (default: ctx->stream)
    std::string err;
    int64_t launches = 0;
    int numSMs = 148;
    // optional per-kernel timing (CUDA events on the launching stream)
    bool timing = false;
    struct Ev { cudaEvent_t a, b; int kind; };
    std::vector<Ev> pending;
    std::vector<cudaEvent_t> evpool;
    double kms[PAC_NKINDS] = {0};
    int64_t kcount[PAC_NKINDS] = {0};
    // constant tables per N (the context's own N plus any N the L2 entry points were asked for)
    std::map<int, TableSet<float>> tf;
    std::map<int, TableSet<double>> td;
    std::map<int, void *> winTables;      // key = kind*65536 + log2N
    std::map<int, std::pair<FastTables, void *>> fastTables;
    // Huffman
    unsigned long long *lenLut = nullptr;
    ulonglong4 *lenLut4 = nullptr;
    uint32_t *codeLen = nullptr;
    uint32_t *decLut = nullptr;
    int32_t *trieChild = nullptr, *trieSym = nullptr;
    DecodeTables dt{};
    // workspaces
    DBuf w_pcm, w_out, w_ns, w_state, w_lines, w_smr, w_bmax, w_osc, w_lrms, w_ba, w_sf, w_tid, w_nby, w_coff,
        w_trE, w_trD, w_hdr, w_ovf, w_dbg1, w_dbg2, w_misc, w_misc2, w_misc3, w_misc4, w_misc5, w_misc6, w_toff;
};

#define CK(call)                                                                                  \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) {                                                                  \
            char b_[512];                                                                         \
            snprintf(b_, sizeof b_, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            ctx->err = b_;                                                                        \
            return PAC_E_CUDA;                                                                    \
        }                                                                                         \
    } while (0)

#define FAIL(code, ...)                                   \
    do {                                                  \
        char b_[512];                                     \
        snprintf(b_, sizeof b_, __VA_ARGS__);             \
        ctx->err = b_;                                    \
        return (code);                                    \
    } while (0)


// ------------------------------------------------------------------ per-kernel timing
static cudaEvent_t ev_get(PacCtx *ctx) {
    if (!ctx->evpool.empty()) { cudaEvent_t e = ctx->evpool.back(); ctx->evpool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
static void timing_flush(PacCtx *ctx) {      // call only after the stream has been synchronised
    for (auto &p : ctx->pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { ctx->kms[p.kind] += ms; ctx->kcount[p.kind]++; }
        ctx->evpool.push_back(p.a); ctx->evpool.push_back(p.b);
    }
    ctx->pending.clear();
}
struct KTimer {
    PacCtx *ctx; int kind; cudaEvent_t a{}, b{};
    cudaStream_t st;
    KTimer(PacCtx *c, int k) : ctx(c), kind(k), st(c->launchStream ? c->launchStream : c->stream) { if (ctx->timing) { a = ev_get(ctx); b = ev_get(ctx); cudaEventRecord(a, st); } }
    ~KTimer() { if (ctx->timing) { cudaEventRecord(b, st); ctx->pending.push_back({a, b, kind}); } }
};

extern "C" int pac_set_stream(PacCtx *ctx, void *stream) {
    if (!ctx) return PAC_E_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    timing_flush(ctx);
    // NULL is CUDA's legacy default stream (what torch.cuda.current_stream().cuda_stream reports for torch's default
    // stream): cudaStreamLegacy names it explicitly, so the internal non-blocking streams are fenced against it by events
    // like against any other caller stream.  The context's private stream has its own sentinel.
    if (stream == PAC_STREAM_OWN) ctx->stream = ctx->ownStream;
    else ctx->stream = stream ? (cudaStream_t)stream : cudaStreamLegacy;
    return PAC_OK;
}

extern "C" int pac_timing_enable(PacCtx *ctx, int on) {
    if (!ctx) return PAC_E_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    timing_flush(ctx);
    ctx->timing = on != 0;
    for (int i = 0; i < PAC_NKINDS; i++) { ctx->kms[i] = 0; ctx->kcount[i] = 0; }
    return PAC_OK;
}
extern "C" int pac_timing_get(PacCtx *ctx, double *ms, int64_t *count) {
    if (!ctx || !ms || !count) return PAC_E_ARG;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    timing_flush(ctx);
    for (int i = 0; i < PAC_NKINDS; i++) { ms[i] = ctx->kms[i]; count[i] = ctx->kcount[i]; }
    return PAC_OK;
}

static inline cudaStream_t LS(PacCtx *ctx) { return ctx->launchStream ? ctx->launchStream : ctx->stream; }

static int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

// ------------------------------------------------------------------ reference constants, computed in double
// psychoac.py:56-64
static double h_bark(double f) { double khz = f / 1000.0, t = khz / 7.5; return 13.0 * atan(khz * 0.76) + 3.5 * atan(t * t); }
// psychoac.py:44-54
static double h_thresh(double f) {
    if (f < 10) f = 10;
    double khz = f / 1000.0;
    return 3.64 * pow(khz, -0.8) - 6.5 * exp(-0.6 * ((khz - 3.3) * (khz - 3.3))) + 0.001 * pow(khz, 4);
}

// psychoac.py:122-156 (cbFreqLimits, AssignMDCTLinesFromFreqLimits)
static void h_band_layout(int nMDCTLines, int sampleRate, int32_t *nLines) {
    static const double lim[25] = {100.0, 200.0, 300.0, 400.0, 510.0, 630.0, 770.0, 920.0, 1080.0, 1270.0, 1480.0, 1720.0, 2000.0,
                                   2320.0, 2700.0, 3150.0, 3700.0, 4400.0, 5300.0, 6400.0, 7700.0, 9500.0, 12000.0, 15500.0, 24000.0};
    double nyq_int = (double)(sampleRate / 2);      // `sampleRate / 2` with an int sampleRate under Python 2 (:133)
    double lower = 0.0;
    for (int b = 0; b < 25; b++) {
        double upper = lim[b] >= sampleRate / 2.0 ? sampleRate / 2.0 : lim[b];
        int cnt = 0;
        for (int i = 0; i < nMDCTLines; i++) {
            double f = (i + 0.5) / nMDCTLines * nyq_int;
            if (f <= upper && f > lower) cnt++;
        }
        nLines[b] = cnt;
        lower = upper;
    }
}

// window.py:56-78 with alpha = 4: Kaiser points over N/2 + 1 samples, cumulative power of the first N/2, mirrored
static void kbd_window_host(int N, double *w) {
    auto i0 = [](double x) { double s = 1, t = 1, q = x * x / 4; for (int k = 1; k < 500; k++) { t *= q / ((double)k * k); s += t; if (t < s * 1e-18) break; } return s; };
    const double Nf = (double)N;
    const int half = N / 2;
    std::vector<double> kz(half + 1);
    double den = 0, d0 = i0(M_PI * 4.0);
    for (int t = 0; t <= half; t++) { double u = 4.0 * t / Nf - 1.0, a = 1.0 - u * u; if (a < 0) a = 0; kz[t] = i0(4.0 * M_PI * sqrt(a)) / d0; den += kz[t] * kz[t]; }
    double c = 0;
    for (int t = 0; t < half; t++) { c += kz[t] * kz[t]; w[t] = w[N - 1 - t] = sqrt(c / den); }
}

template <typename T>
static int build_tables(PacCtx *ctx, int N, TableSet<T> &ts) {
    using T2 = typename Vec2<T>::type;
    const int M = N / 2, H = M / 2, fs = ctx->p.sampleRate;
    size_t nT = (size_t)N * 2 + (size_t)M * 4;
    size_t nT2 = (size_t)M + (M + 1) + H + H;
    size_t bytes = nT * sizeof(T) + nT2 * sizeof(T2) + M + 64;
    std::vector<unsigned char> host(bytes);
    T2 *t2 = reinterpret_cast<T2 *>(host.data());
    T2 *tw = t2, *tws = tw + M, *pre = tws + (M + 1), *post = pre + H;
    T *t1 = reinterpret_cast<T *>(post + H);
    T *sinw = t1, *hann = sinw + N, *zline = hann + N, *tiq = zline + M, *mld = tiq + M, *zpeak = mld + M;
    uint8_t *bol = reinterpret_cast<uint8_t *>(zpeak + M);
    for (int m = 0; m < M; m++) { double a = -2.0 * M_PI * m / M; tw[m] = mk2<T>((T)cos(a), (T)sin(a)); }
    for (int k = 0; k <= M; k++) { double a = -2.0 * M_PI * k / N; tws[k] = mk2<T>((T)cos(a), (T)sin(a)); }
    for (int n = 0; n < H; n++) {
        double a = -M_PI * n / M, b = -M_PI * (n + 0.25) / M;
        pre[n] = mk2<T>((T)cos(a), (T)sin(a));
        post[n] = mk2<T>((T)cos(b), (T)sin(b));
    }
    double Nf = (double)N;
    for (int n = 0; n < N; n++) {
        sinw[n] = (T)sin((n + 0.5) * M_PI / Nf);                                   // window.py:35-37
        hann[n] = (T)(0.5 * (1 - cos(2.0 * (n + 0.5) * M_PI / Nf)));               // window.py:49-51
    }
    if (ctx->p.window == PAC_WINDOW_KBD && N == ctx->N) {                          // window.py:56-78, alpha = 4 (the codec's own block size only)
        std::vector<double> kw(N);
        kbd_window_host(N, kw.data());
        for (int n = 0; n < N; n++) sinw[n] = (T)kw[n];
    }
    ts.dev.winInPlace = (ctx->p.window == PAC_WINDOW_KBD && N == ctx->N) ? 0 : 1;
    double mx = 0;
    std::vector<double> mldd(M);
    for (int i = 0; i < M; i++) {
        double f = fs / 2.0 / M * (i + 0.5);                                       // psychoac.py:434
        zline[i] = (T)h_bark(f);
        tiq[i] = (T)pow(10.0, (h_thresh(f) - 96) / 10);                            // :437
        double f2 = ((i + 0.5) / M) * (fs / 2.0);                                  // :570
        mldd[i] = pow(10.0, 1.25 * (1 - cos(M_PI * (fmin(f2, 3000.) / 3000.)) - 2.5));   // :367
        if (mldd[i] > mx) mx = mldd[i];
        zpeak[i] = (T)h_bark((double)i * (double)(fs / N));                        // :186-188, integer fs/N
    }
    for (int i = 0; i < M; i++) mld[i] = (T)(mldd[i] / mx);                        // :370
    memset(bol, 0, M);
    if (M == ctx->M)
        for (int b = 0; b < ctx->bands.nBands; b++)
            for (int i = ctx->bands.lo[b]; i < ctx->bands.lo[b + 1]; i++) bol[i] = (uint8_t)b;
    CK(cudaMalloc(&ts.mem, bytes));
    CK(cudaMemcpy(ts.mem, host.data(), bytes, cudaMemcpyHostToDevice));
    unsigned char *d = reinterpret_cast<unsigned char *>(ts.mem);
    auto dv = [&](const void *hp) { return d + (reinterpret_cast<const unsigned char *>(hp) - host.data()); };
    ts.dev.tw = reinterpret_cast<const T2 *>(dv(tw));
    ts.dev.tw_split = reinterpret_cast<const T2 *>(dv(tws));
    ts.dev.mdct_pre = reinterpret_cast<const T2 *>(dv(pre));
    ts.dev.mdct_post = reinterpret_cast<const T2 *>(dv(post));
    ts.dev.sinw = reinterpret_cast<const T *>(dv(sinw));
    ts.dev.hann = reinterpret_cast<const T *>(dv(hann));
    ts.dev.zline = reinterpret_cast<const T *>(dv(zline));
    ts.dev.tiq = reinterpret_cast<const T *>(dv(tiq));
    ts.dev.mld = reinterpret_cast<const T *>(dv(mld));
    ts.dev.zpeak = reinterpret_cast<const T *>(dv(zpeak));
    ts.dev.band_of_line = reinterpret_cast<const uint8_t *>(dv(bol));
    ts.dev.hann_w = mk2<T>((T)cos(M_PI / N), (T)sin(M_PI / N));
    ts.dev.cnorm = (T)(8.0 / 3.0 * 4.0 / ((double)N * (double)N));                 // psychoac.py:448
    ts.dev.imdct_scale = (T)2;
    return PAC_OK;
}


// ------------------------------------------------------------------ fp32 fast-path geometry (tests/model_analysis.py: Geometry)
struct FastSet { FastTables dev; void *mem = nullptr; };

static int build_fast_tables(PacCtx *ctx, int N, FastSet &fsd) {
    const int M = N / 2, NT = M / 4, NW = NT / 32, fs = ctx->p.sampleRate;
    std::vector<double> zl(M), zp(M);
    for (int i = 0; i < M; i++) {
        zl[i] = h_bark(fs / 2.0 / M * (i + 0.5));
        zp[i] = h_bark((double)i * (double)(fs / N));
    }
    const double dn = -27.0 * (log2(10.0) / 10.0);
    std::vector<int> eL(M), eU(M);
    for (int k = 0; k < M; k++) {
        int el = -1, eu = M;
        for (int i = 0; i < M; i++) {
            double dz = zl[i] - zp[k];
            if (dz < -0.5) el = i;                         // |dz| > .5 exactly as psychoac.py:116
            if (dz > 0.5 && eu == M) eu = i;
        }
        eL[k] = el; eU[k] = eu;
    }
    for (int k = 1; k < M; k++)
        if (eL[k] < eL[k - 1] || eU[k] < eU[k - 1] || zp[k] < zp[k - 1]) FAIL(PAC_E_ARG, "Bark tables are not monotone");
    if (M > 32767) FAIL(PAC_E_ARG, "fast tables pack indices in 16 bits");
    // per line: plateau window [pa, pb) = bins with eL[k] < i < eU[k]; bins below pa reach the line with their upper skirt, bins
    // from pb on with their lower skirt.  Static factors from the scan entry to the line (exponents <= 0).
    std::vector<float> lineRec(4 * (size_t)M), lineZ(2 * (size_t)M), binZ(2 * (size_t)M);
    std::vector<short> kcountU(M + 1);
    for (int i = 0; i < M; i++) {
        int pa = 0, pb = 0;
        for (int k = 0; k < M; k++) {
            if (eU[k] <= i) pa = k + 1;
            if (eL[k] < i) pb = k + 1;
        }
        const uint32_t pk = (uint32_t)pa | ((uint32_t)pb << 16);
        float fL = 0.f, fU = 0.f;
        if (pb < M) { const double e = dn * (zp[pb] - 0.5 - zl[i]); if (e > 1e-9) FAIL(PAC_E_ARG, "lower-skirt factor > 1"); fL = (float)exp2(e); }
        if (pa > 0) { const double e = dn * (zl[i] - zp[pa - 1] - 0.5); if (e > 1e-9) FAIL(PAC_E_ARG, "upper-skirt factor > 1"); fU = (float)exp2(e); }
        memcpy(&lineRec[4 * (size_t)i], &pk, 4);
        lineRec[4 * (size_t)i + 1] = fL; lineRec[4 * (size_t)i + 2] = fU;
        lineRec[4 * (size_t)i + 3] = (float)pow(10.0, (h_thresh(fs / 2.0 / M * (i + 0.5)) - 96) / 10);      // psychoac.py:437
        kcountU[i] = (short)pa;                            // bins with eU <= i
        const float lh = (float)zl[i];
        lineZ[2 * (size_t)i] = lh; lineZ[2 * (size_t)i + 1] = (float)(zl[i] - (double)lh);
        const float zh = (float)zp[i];
        binZ[2 * (size_t)i] = zh; binZ[2 * (size_t)i + 1] = (float)(zp[i] - (double)zh);
    }
    kcountU[M] = (short)M;
    std::vector<unsigned short> binEU(M);
    for (int k = 0; k < M; k++) binEU[k] = (unsigned short)eU[k];
    // scan weights between bins
    auto om = [&](int i, int j) -> float {
        if (i < 0 || j < 0 || i >= M || j >= M) return 0.f;
        return (float)exp2(dn * fabs(zp[j] - zp[i]));
    };
    std::vector<float> sD(13 * (size_t)NT), sA(10 * (size_t)NT);
    for (int vt = 0; vt < NT; vt++) {
        const int lane = vt & 31, chunk = vt >> 5, b = 4 * vt;
        for (int q = 0; q < 3; q++) sD[q * NT + vt] = om(b + q, b + q + 1);
        for (int s = 0; s < 5; s++) sD[(3 + s) * NT + vt] = (lane + (1 << s) < 32) ? om(b, 4 * (vt + (1 << s))) : 0.f;
        sD[8 * NT + vt] = (chunk + 1 < NW) ? om(4 * (vt + 1) < M ? 4 * (vt + 1) : M - 1, 128 * (chunk + 1)) : 0.f;
        if (lane == 31 && chunk + 1 < NW) sD[8 * NT + vt] = 1.f;
        for (int q = 0; q < 4; q++) sD[(9 + q) * NT + vt] = (b + 4 < M) ? om(b + q, b + 4) : 0.f;
        for (int s = 0; s < 5; s++) sA[s * NT + vt] = (lane >= (1 << s)) ? om(4 * (vt - (1 << s)) + 3, b + 3) : 0.f;
        sA[5 * NT + vt] = chunk > 0 ? (lane == 0 ? 1.f : om(128 * chunk - 1, b - 1)) : 0.f;
        for (int q = 0; q < 4; q++) sA[(6 + q) * NT + vt] = b > 0 ? om(b - 1, b + q) : 0.f;
    }
    for (int c = 0; c < 16; c++) {
        fsd.dev.omD[c] = (c + 1 < NW) ? om(128 * c, 128 * (c + 1)) : 0.f;
        fsd.dev.omA[c] = (c >= 1 && c < NW) ? om(128 * c - 1, 128 * c + 127) : 0.f;
    }
    size_t bytes = lineRec.size() * 4 + lineZ.size() * 4 + binZ.size() * 4 + (size_t)(M + 1) * 2 + (size_t)M * 2 + (sD.size() + sA.size()) * 4 + 512;
    std::vector<unsigned char> host(bytes, 0);
    size_t o = 0;
    auto put = [&](const void *src, size_t n) { o = (o + 15) & ~(size_t)15; memcpy(host.data() + o, src, n); size_t r = o; o += n; return r; };
    size_t o_lr = put(lineRec.data(), lineRec.size() * 4), o_lz = put(lineZ.data(), lineZ.size() * 4), o_bz = put(binZ.data(), binZ.size() * 4);
    size_t o_kc = put(kcountU.data(), (size_t)(M + 1) * 2), o_be = put(binEU.data(), (size_t)M * 2);
    size_t o_sD = put(sD.data(), sD.size() * 4), o_sA = put(sA.data(), sA.size() * 4);
    CK(cudaMalloc(&fsd.mem, o + 16));
    CK(cudaMemcpy(fsd.mem, host.data(), o, cudaMemcpyHostToDevice));
    unsigned char *d = reinterpret_cast<unsigned char *>(fsd.mem);
    fsd.dev.lineRec = reinterpret_cast<const float4 *>(d + o_lr); fsd.dev.lineZ = reinterpret_cast<const float2 *>(d + o_lz);
    fsd.dev.binZ = reinterpret_cast<const float2 *>(d + o_bz); fsd.dev.kcountU = reinterpret_cast<const short *>(d + o_kc);
    for (int h = 0; h < 32; h++) fsd.dev.kUhc[h] = 64 * h + 63 < M ? ((uint32_t)(unsigned short)kcountU[64 * h] | (uint32_t)(unsigned short)kcountU[64 * h + 63] << 16) : 0u;
    fsd.dev.binEU = reinterpret_cast<const uint2 *>(d + o_be);
    fsd.dev.sD = reinterpret_cast<const float *>(d + o_sD); fsd.dev.sA = reinterpret_cast<const float *>(d + o_sA);
    return PAC_OK;
}

static int get_fast_tables(PacCtx *ctx, int N, FastTables *out);

template <typename T> static std::map<int, TableSet<T>> &tabmap(PacCtx *ctx);
template <> std::map<int, TableSet<float>> &tabmap<float>(PacCtx *ctx) { return ctx->tf; }
template <> std::map<int, TableSet<double>> &tabmap<double>(PacCtx *ctx) { return ctx->td; }

template <typename T>
static int get_tables(PacCtx *ctx, int N, DevTables<T> *out) {
    auto &m = tabmap<T>(ctx);
    auto it = m.find(N);
    if (it == m.end()) {
        TableSet<T> ts;
        int rc = build_tables<T>(ctx, N, ts);
        if (rc) return rc;
        it = m.emplace(N, ts).first;
    }
    *out = it->second.dev;
    return PAC_OK;
}

static int get_fast_tables(PacCtx *ctx, int N, FastTables *out) {
    auto it = ctx->fastTables.find(N);
    if (it == ctx->fastTables.end()) {
        FastSet fsd;
        int rc = build_fast_tables(ctx, N, fsd);
        if (rc) return rc;
        it = ctx->fastTables.emplace(N, std::make_pair(fsd.dev, fsd.mem)).first;
    }
    *out = it->second.first;
    return PAC_OK;
}

// ------------------------------------------------------------------ Huffman tables -> device LUTs
struct HostTrie {
    std::vector<int32_t> child, sym;
    int add() { child.push_back(-1); child.push_back(-1); sym.push_back(-2); return (int)sym.size() - 1; }
};

static int build_huffman(PacCtx *ctx, const PacHuffTables *h) {
    int total = 0;
    for (int t = 0; t < kNTables; t++) {
        if (h->nkeys[t] <= 0 || h->nkeys[t] > kLenLutSize || h->esc_len[t] <= 0 || h->esc_len[t] > 24) FAIL(PAC_E_ARG, "bad Huffman table %d", t + 1);
        ctx->ec.nkeys[t] = h->nkeys[t]; ctx->ec.off[t] = h->off[t];
        ctx->ec.esc_len[t] = h->esc_len[t]; ctx->ec.esc_code[t] = h->esc_code[t];
        if (h->off[t] + h->nkeys[t] > total) total = h->off[t] + h->nkeys[t];
    }
    std::vector<unsigned long long> lut(kLenLutSize, 0ull);
    for (int t = 0; t < kNTables; t++)
        for (int v = 0; v < h->nkeys[t]; v++) {
            unsigned l = h->len[h->off[t] + v];
            if (l > 31) FAIL(PAC_E_ARG, "Huffman code longer than 31 bits");
            lut[v] |= (unsigned long long)l << (5 * t);
        }
    {   // packed-slot variant used by the scan kernel
        std::vector<unsigned long long> l4((size_t)(kLenLutSize + 1) * 4, 0ull);
        for (int v = 0; v <= kLenLutSize; v++)
            for (int t = 0; t < kNTables; t++) {
                unsigned l = (v < kLenLutSize && v < h->nkeys[t]) ? h->len[h->off[t] + v] : 0;
                int w = t < 5 ? 0 : 1, sh = 12 * (t % 5);
                if (l) l4[(size_t)v * 4 + w] |= (unsigned long long)l << sh;
                else { l4[(size_t)v * 4 + w] |= (unsigned long long)h->esc_len[t] << sh; l4[(size_t)v * 4 + 2 + w] |= 1ull << sh; }
            }
        CK(cudaMalloc(&ctx->lenLut4, l4.size() * 8));
        CK(cudaMemcpy(ctx->lenLut4, l4.data(), l4.size() * 8, cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&ctx->lenLut, lut.size() * 8));
    CK(cudaMemcpy(ctx->lenLut, lut.data(), lut.size() * 8, cudaMemcpyHostToDevice));
    {   // the pack kernel's table: code << 5 | length in one word (codes are at most 26 bits: build_huffman rejects lengths > 31, and a
        // code longer than 27 bits could not share the word)
        std::vector<uint32_t> cl((size_t)total);
        for (int i = 0; i < total; i++) {
            if (h->len[i] > 27) FAIL(PAC_E_ARG, "Huffman code longer than 27 bits");
            cl[i] = (h->code[i] << 5) | h->len[i];
        }
        CK(cudaMalloc(&ctx->codeLen, (size_t)total * 4));
        CK(cudaMemcpy(ctx->codeLen, cl.data(), (size_t)total * 4, cudaMemcpyHostToDevice));
    }
    // decoder: binary tries (Huffman.py:321-344 does a string-prefix search; a trie is the same relation)
    HostTrie tr;
    int root[kNTables];
    auto insert = [&](int r, uint32_t code, int len, int sym) {
        int cur = r;
        for (int i = len - 1; i >= 0; i--) {
            int bit = (code >> i) & 1;
            if (tr.child[2 * cur + bit] < 0) { int c = tr.add(); tr.child[2 * cur + bit] = c; }
            cur = tr.child[2 * cur + bit];
        }
        tr.sym[cur] = sym;
    };
    for (int t = 0; t < kNTables; t++) {
        root[t] = tr.add();
        for (int v = 0; v < h->nkeys[t]; v++)
            if (h->len[h->off[t] + v]) insert(root[t], h->code[h->off[t] + v], h->len[h->off[t] + v], v);
        insert(root[t], h->esc_code[t], h->esc_len[t], -1);
    }
    std::vector<uint32_t> dl((size_t)kNTables << kLutBits);
    for (int t = 0; t < kNTables; t++)
        for (uint32_t pfx = 0; pfx < (1u << kLutBits); pfx++) {
            int cur = root[t];
            uint32_t e = 0;
            bool done = false;
            for (int d = 0; d < kLutBits; d++) {
                if (tr.sym[cur] != -2) { e = kLutLeaf | ((uint32_t)(tr.sym[cur] + 1) << 8) | (uint32_t)d; done = true; break; }
                int bit = (pfx >> (kLutBits - 1 - d)) & 1;
                cur = tr.child[2 * cur + bit];
                if (cur < 0) { e = kLutInvalid; done = true; break; }
            }
            if (!done) e = tr.sym[cur] != -2 ? (kLutLeaf | ((uint32_t)(tr.sym[cur] + 1) << 8) | (uint32_t)kLutBits) : (uint32_t)cur;
            dl[((size_t)t << kLutBits) + pfx] = e;
        }
    CK(cudaMalloc(&ctx->decLut, dl.size() * 4));
    CK(cudaMemcpy(ctx->decLut, dl.data(), dl.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&ctx->trieChild, tr.child.size() * 4));
    CK(cudaMemcpy(ctx->trieChild, tr.child.data(), tr.child.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&ctx->trieSym, tr.sym.size() * 4));
    CK(cudaMemcpy(ctx->trieSym, tr.sym.data(), tr.sym.size() * 4, cudaMemcpyHostToDevice));
    ctx->dt.lut = ctx->decLut; ctx->dt.child = ctx->trieChild; ctx->dt.sym = ctx->trieSym;
    for (int t = 0; t < kNTables; t++) ctx->dt.root[t] = root[t];
    return PAC_OK;
}

// ------------------------------------------------------------------ context
extern "C" const char *pac_version(void) { return "pac-b200 0.1 (sm_100a)"; }

extern "C" const char *pac_last_error(PacCtx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int64_t pac_launch_count(PacCtx *ctx) { return ctx ? ctx->launches : 0; }

static int ctx_init(PacCtx *ctx, int device, int precision, const PacParams *params, const PacHuffTables *tables) {
    if (!params || !tables) FAIL(PAC_E_ARG, "null params/tables");
    if (precision != PAC_PRECISION_FP64 && precision != PAC_PRECISION_FP32) FAIL(PAC_E_ARG, "precision must be 0 (fp64) or 1 (fp32)");
    if (params->nChannels != 2) FAIL(PAC_E_ARG, "only 2 channels are supported (as in the reference, codec.py:46-47)");
    if (params->nMDCTLines != 1024 && params->nMDCTLines != 512) FAIL(PAC_E_ARG, "nMDCTLines must be 1024 or 512");
    if (params->nScaleBits < 1 || params->nScaleBits > 4 || params->nMantSizeBits < 1 || params->nMantSizeBits > 4)
        FAIL(PAC_E_ARG, "unsupported bit-field widths");
    // the container does not store nTableIDBits and the reference's reader hard-codes 4 (pacfile.py:189): any other width
    // would write images that neither this library nor the reference can decode
    if (params->nTableIDBits != 4) FAIL(PAC_E_ARG, "nTableIDBits must be 4 (the .pac reader hard-codes it, pacfile.py:189)");
    if (params->sampleRate < 8000 || params->sampleRate > 192000) FAIL(PAC_E_ARG, "unsupported sample rate");
    if (params->window != PAC_WINDOW_SINE && params->window != PAC_WINDOW_KBD) FAIL(PAC_E_ARG, "window must be PAC_WINDOW_SINE or PAC_WINDOW_KBD");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) FAIL(PAC_E_NODEVICE, "no CUDA device: the engine has no CPU fallback");
    if (device < 0 || device >= ndev) FAIL(PAC_E_ARG, "device %d out of range (%d devices)", device, ndev);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) FAIL(PAC_E_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    ctx->numSMs = prop.multiProcessorCount;
    ctx->device = device; ctx->precision = precision; ctx->p = *params;
    ctx->M = params->nMDCTLines; ctx->N = 2 * ctx->M; ctx->LOGM = ilog2(ctx->M);
    h_band_layout(ctx->M, params->sampleRate, ctx->nLines);
    ctx->bands.nBands = 25;
    int lo = 0;
    for (int b = 0; b < 25; b++) { ctx->bands.lo[b] = (int16_t)lo; lo += ctx->nLines[b]; }
    for (int b = 25; b <= kMaxBands; b++) ctx->bands.lo[b] = (int16_t)lo;
    if (lo != ctx->M) FAIL(PAC_E_ARG, "band layout does not cover the MDCT lines (%d of %d)", lo, ctx->M);
    EncConsts &ec = ctx->ec;
    const int NB = ctx->bands.nBands;
    double bb = params->targetBitsPerSample * ctx->M;     // codec.py:223
    bb -= params->nScaleBits * (NB + 1);                  // :224
    bb -= params->nMantSizeBits * NB;                     // :225
    bb -= params->nTableIDBits;                           // :227
    ec.bitBudget = bb;
    ec.nScaleBits = params->nScaleBits; ec.nMantSizeBits = params->nMantSizeBits; ec.nTableIDBits = params->nTableIDBits;
    ec.maxMantBits = (1 << params->nMantSizeBits) > 16 ? 16 : (1 << params->nMantSizeBits);     // codec.py:218-219
    ec.fixedBits = params->nScaleBits + params->nTableIDBits + NB * (params->nMantSizeBits + params->nScaleBits) + NB;
    CK(cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking));
    ctx->stream = ctx->ownStream;
    {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CK(cudaStreamCreateWithPriority(&ctx->sA, cudaStreamNonBlocking, lo));      // analysis: lowest priority
        CK(cudaStreamCreateWithPriority(&ctx->sB, cudaStreamNonBlocking, hi));      // scan + pack: highest
        CK(cudaStreamCreateWithPriority(&ctx->sM, cudaStreamNonBlocking, lo));      // MDCT of the tiles ahead: lowest
        CK(cudaStreamCreateWithPriority(&ctx->sD, cudaStreamNonBlocking, hi));      // drain: a handful of warps, placed promptly
        CK(cudaEventCreateWithFlags(&ctx->evStart, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&ctx->sC, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) { CK(cudaEventCreateWithFlags(&ctx->evH[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&ctx->evD[i], cudaEventDisableTiming)); }
    }
    int rc = build_huffman(ctx, tables);
    if (rc) return rc;
    DevTables<float> tf; DevTables<double> td;
    if ((rc = get_tables<float>(ctx, ctx->N, &tf))) return rc;
    if ((rc = get_tables<double>(ctx, ctx->N, &td))) return rc;
    return PAC_OK;
}

extern "C" int pac_ctx_create(int device, int precision, const PacParams *params, const PacHuffTables *tables, PacCtx **out) {
    if (!out) return PAC_E_ARG;
    *out = nullptr;
    PacCtx *ctx = new PacCtx();
    int rc = ctx_init(ctx, device, precision, params, tables);
    if (rc) { g_create_error = ctx->err; pac_ctx_destroy(ctx); return rc; }
    *out = ctx;
    return PAC_OK;
}

extern "C" void pac_ctx_destroy(PacCtx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) { cudaStreamSynchronize(ctx->stream); timing_flush(ctx); }
    if (ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
    if (ctx->sA) cudaStreamDestroy(ctx->sA);
    if (ctx->sB) cudaStreamDestroy(ctx->sB);
    if (ctx->sC) cudaStreamDestroy(ctx->sC);
    if (ctx->sM) cudaStreamDestroy(ctx->sM);
    if (ctx->sD) cudaStreamDestroy(ctx->sD);
    for (int i = 0; i < 2; i++) { if (ctx->evH[i]) cudaEventDestroy(ctx->evH[i]); if (ctx->evD[i]) cudaEventDestroy(ctx->evD[i]); }
    ctx->w_pcm2.release(); ctx->w_out2.release();
    if (ctx->evStart) cudaEventDestroy(ctx->evStart);
    for (cudaEvent_t e : ctx->evA) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->evB) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->evM) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->evG) cudaEventDestroy(e);
    for (int i = 0; i < 2; i++) for (cudaEvent_t e : ctx->evS[i]) cudaEventDestroy(e);
    ctx->hostState.release();
    for (cudaEvent_t e : ctx->evpool) cudaEventDestroy(e);
    for (auto &kv : ctx->tf) cudaFree(kv.second.mem);
    for (auto &kv : ctx->td) cudaFree(kv.second.mem);
    for (auto &kv : ctx->winTables) cudaFree(kv.second);
    for (auto &kv : ctx->fastTables) cudaFree(kv.second.second);
    cudaFree(ctx->lenLut); cudaFree(ctx->lenLut4); cudaFree(ctx->codeLen);
    cudaFree(ctx->decLut); cudaFree(ctx->trieChild); cudaFree(ctx->trieSym);
    DBuf *bufs[] = {&ctx->w_pcm, &ctx->w_out, &ctx->w_ns, &ctx->w_state, &ctx->w_lines, &ctx->w_smr, &ctx->w_bmax, &ctx->w_osc,
                    &ctx->w_lrms, &ctx->w_ba, &ctx->w_sf, &ctx->w_tid, &ctx->w_nby, &ctx->w_coff, &ctx->w_trE, &ctx->w_trD,
                    &ctx->w_hdr, &ctx->w_ovf, &ctx->w_dbg1, &ctx->w_dbg2, &ctx->w_misc, &ctx->w_misc2, &ctx->w_misc3,
                    &ctx->w_misc4, &ctx->w_misc5, &ctx->w_misc6, &ctx->w_toff};
    for (DBuf *b : bufs) b->release();
    delete ctx;
}

extern "C" void *pac_pinned_alloc(size_t nbytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, nbytes ? nbytes : 1, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void pac_pinned_free(void *p) { if (p) cudaFreeHost(p); }

// host -> device copy bandwidth of a (pinned) host buffer on `device`, GB/s: lets a caller that cannot learn the GPU's NUMA node from
// the OS pick the host memory placement by measurement (bench.py's e2e leg)
extern "C" int pac_h2d_bandwidth(const void *host, size_t nbytes, int device, double *gbs) {
    if (!host || !nbytes || !gbs) return PAC_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return PAC_E_CUDA; }
    void *d = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = PAC_E_CUDA;
    float ms = 0.f;
    if (cudaMalloc(&d, nbytes) == cudaSuccess && cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess &&
        cudaMemcpy(d, host, nbytes, cudaMemcpyHostToDevice) == cudaSuccess &&            // warm
        cudaEventRecord(e0, 0) == cudaSuccess && cudaMemcpyAsync(d, host, nbytes, cudaMemcpyHostToDevice, 0) == cudaSuccess &&
        cudaEventRecord(e1, 0) == cudaSuccess && cudaEventSynchronize(e1) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess && ms > 0.f) {
        *gbs = (double)nbytes / ((double)ms * 1e-3) / 1e9;
        rc = PAC_OK;
    }
    if (rc != PAC_OK) cudaGetLastError();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (d) cudaFree(d);
    return rc;
}

extern "C" int pac_band_layout(PacCtx *ctx, int32_t *nLines, int32_t *nBands) {
    if (!ctx || !nLines || !nBands) return PAC_E_ARG;
    for (int b = 0; b < ctx->bands.nBands; b++) nLines[b] = ctx->nLines[b];
    *nBands = ctx->bands.nBands;
    return PAC_OK;
}

extern "C" int64_t pac_num_blocks(PacCtx *ctx, int64_t nSamples) {
    if (!ctx || nSamples < 0) return PAC_E_ARG;
    return (nSamples + ctx->M - 1) / ctx->M + 1;
}

static int header_bytes(PacCtx *ctx) { return 4 + 18 + 4 + 2 * ctx->bands.nBands; }

extern "C" int64_t pac_encode_bound(PacCtx *ctx, int64_t nSamples) {
    if (!ctx || nSamples < 0) return PAC_E_ARG;
    // target rate + a generous reservoir allowance; K5 refuses to write past the capacity it is given.
    int64_t nb = pac_num_blocks(ctx, nSamples);
    double perCh = ctx->p.targetBitsPerSample * ctx->M * 1.5 + 2048;
    int64_t perBlock = 2 * (4 + (int64_t)(perCh / 8) + 1);
    return header_bytes(ctx) + nb * perBlock + 4096;
}

static bool is_device_ptr(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// device-visible alias of a pinned (page-locked, mapped) HOST allocation, or NULL: kernels can then write results straight into
// the caller's host buffer across PCIe instead of into a device staging buffer that is copied back afterwards
static void *mapped_host_alias(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type == cudaMemoryTypeHost && at.devicePointer) return at.devicePointer;
    return nullptr;
}

// pacfile.py:237-261
static void build_header(PacCtx *ctx, int64_t nSamples, uint8_t *h) {
    auto le32 = [](uint8_t *p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; };
    auto le16 = [](uint8_t *p, uint32_t v) { p[0] = v; p[1] = v >> 8; };
    memcpy(h, "PAC ", 4);
    uint32_t ns = (uint32_t)nSamples;
    if (nSamples % ctx->M == 0) ns += ctx->M;            // :240-242 (the padding rule as written)
    le32(h + 4, (uint32_t)ctx->p.sampleRate);
    le16(h + 8, (uint32_t)ctx->p.nChannels);
    le32(h + 10, ns);
    le32(h + 14, (uint32_t)ctx->M);
    le16(h + 18, (uint32_t)ctx->p.nScaleBits);
    le16(h + 20, (uint32_t)ctx->p.nMantSizeBits);
    le32(h + 22, (uint32_t)ctx->bands.nBands);
    for (int b = 0; b < ctx->bands.nBands; b++) le16(h + 26 + 2 * b, (uint32_t)ctx->nLines[b]);
}

// ------------------------------------------------------------------ kernel launch helpers
template <typename T, int LOGM>
static int launch_analysis_t(PacCtx *ctx, AnalysisArgs<T> &a) {
    size_t smem = sizeof(AnalysisSmem<T, LOGM, sizeof(T) == 4>);
    if (const char *ex = getenv("PAC_EXTRA_SMEM")) smem += (size_t)atoi(ex);      // occupancy experiments
    a.poisonOn = 0; a.poison = 0; a.smemWords = (uint32_t)(sizeof(AnalysisSmem<T, LOGM, sizeof(T) == 4>) / 4);
    if (const char *po = getenv("PAC_POISON_SMEM")) { a.poisonOn = 1; a.poison = (uint32_t)strtoul(po, nullptr, 16); }
    CK(cudaFuncSetAttribute(k_analysis<T, LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (const char *cv = getenv("PAC_CARVEOUT_KB"))                                // L1-size experiments
        CK(cudaFuncSetAttribute(k_analysis<T, LOGM>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(cv) * 100 / 228));
    int perSM = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_analysis<T, LOGM>, (1 << LOGM) / 4, smem));
    if (perSM < 1) perSM = 1;
    int64_t grid = (int64_t)ctx->numSMs * perSM;
    if (ctx->launchStream) {
        // overlapped tiles: short-lived CTAs let the scan/pack kernels of the previous tile in (a persistent grid measured 460 vs
        // 354 ms: it never yields SMs).  A few blocks per CTA amortise the per-thread table prologue (4 % of a one-block CTA).
        static const int bpc = getenv("PAC_BLOCKS_PER_CTA") ? atoi(getenv("PAC_BLOCKS_PER_CTA")) : 4;
        const int64_t g2 = (a.nwork + (bpc > 0 ? bpc : 1) - 1) / (bpc > 0 ? bpc : 1);
        if (g2 > grid) grid = g2;
    }
    if (grid > a.nwork) grid = a.nwork;
    if (grid < 1) grid = 1;
    { KTimer kt(ctx, PAC_K_ANALYSIS); k_analysis<T, LOGM><<<(unsigned)grid, (1 << LOGM) / 4, smem, LS(ctx)>>>(a); }
    ctx->launches++;
    CK(cudaGetLastError());
    return PAC_OK;
}

// the MDCT-only stage instantiation (sections A + C): persistent grid, the part of the shared-memory struct it touches
template <typename T, int LOGM>
static int launch_mdct_only_t(PacCtx *ctx, AnalysisArgs<T> &a) {
    using SM = AnalysisSmem<T, LOGM, sizeof(T) == 4>;
    const size_t smem = offsetof(SM, fs);
    a.poisonOn = 0; a.poison = 0; a.smemWords = 0;
    CK(cudaFuncSetAttribute(k_analysis<T, LOGM, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int perSM = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_analysis<T, LOGM, true>, (1 << LOGM) / 4, smem));
    if (perSM < 1) perSM = 1;
    int64_t grid = (int64_t)ctx->numSMs * perSM;
    if (grid > a.nwork) grid = a.nwork;
    if (grid < 1) grid = 1;
    { KTimer kt(ctx, PAC_K_MDCT); k_analysis<T, LOGM, true><<<(unsigned)grid, (1 << LOGM) / 4, smem, LS(ctx)>>>(a); }
    ctx->launches++;
    CK(cudaGetLastError());
    return PAC_OK;
}

// K1 of the fp32 fast mode: window + MDCT + overall scale as their own fp64 kernel (mdct.cuh), one CTA per stereo block
template <int LOGM, bool PCM>
static int launch_mdct_t(PacCtx *ctx, const MdctArgs &m) {
    const size_t smem = sizeof(EncMdctSmem<LOGM>);
    CK(cudaFuncSetAttribute(k_mdct_enc<LOGM, PCM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int perSM = 1;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_mdct_enc<LOGM, PCM>, (1 << LOGM) / 8, smem));
    if (perSM < 1) perSM = 1;
    int64_t grid = (int64_t)ctx->numSMs * perSM * 2;          // a few work items per CTA: the twiddle tables are staged once per CTA
    if (grid > m.nwork) grid = m.nwork;
    if (grid < 1) grid = 1;
    { KTimer kt(ctx, PAC_K_MDCT); k_mdct_enc<LOGM, PCM><<<(unsigned)grid, (1 << LOGM) / 8, smem, LS(ctx)>>>(m); }
    ctx->launches++;
    CK(cudaGetLastError());
    return PAC_OK;
}

static int launch_mdct(PacCtx *ctx, const AnalysisArgs<float> &a) {
    MdctArgs m{};
    m.pcm = a.pcm; m.strideSamples = a.strideSamples; m.nSamples = a.nSamples; m.blocks = a.blocks;
    m.S = a.S; m.b0 = a.b0; m.nb = a.nb; m.nwork = a.nwork; m.nScaleBits = ctx->p.nScaleBits;
    m.lines = a.lines; m.oscale = a.oscale;
    int rc = get_tables<double>(ctx, ctx->N, &m.tabd);
    if (rc) return rc;
    if (ctx->LOGM == 10) return m.pcm ? launch_mdct_t<10, true>(ctx, m) : launch_mdct_t<10, false>(ctx, m);
    if (ctx->LOGM == 9) return m.pcm ? launch_mdct_t<9, true>(ctx, m) : launch_mdct_t<9, false>(ctx, m);
    FAIL(PAC_E_ARG, "unsupported nMDCTLines");
}
static int launch_mdct(PacCtx *, const AnalysisArgs<double> &) { return PAC_OK; }     // fp64 mode: the MDCT is a section of k_analysis

template <typename T>
static int launch_analysis(PacCtx *ctx, AnalysisArgs<T> &a, bool withMdct = true) {
    if (withMdct) {   // fp32 mode: k_mdct first, on the same stream; k_analysis picks its lines and scales up from a.lines / a.oscale
        int rcm = launch_mdct(ctx, a);
        if (rcm) return rcm;
    }
    a.bands = ctx->bands;
    a.nScaleBits = ctx->p.nScaleBits;
    int rc = get_tables<T>(ctx, ctx->N, &a.tab);
    if (rc) return rc;
    if ((rc = get_tables<double>(ctx, ctx->N, &a.tabd))) return rc;
    if (sizeof(T) == 4 && (rc = get_fast_tables(ctx, ctx->N, &a.ft))) return rc;
    if (ctx->LOGM == 10) return launch_analysis_t<T, 10>(ctx, a);
    if (ctx->LOGM == 9) return launch_analysis_t<T, 9>(ctx, a);
    FAIL(PAC_E_ARG, "unsupported nMDCTLines");
}

template <typename T>
static int launch_scan(PacCtx *ctx, ScanArgs<T> &a) {
    a.ec = ctx->ec; a.bands = ctx->bands; a.M = ctx->M; a.lenLut = ctx->lenLut; a.lenLut4 = ctx->lenLut4;
    {
        DevTables<T> tb;
        int rc = get_tables<T>(ctx, ctx->N, &tb);
        if (rc) return rc;
        a.band_of_line = tb.band_of_line;
    }
    // Warps per stream.  Many streams: the kernel is a throughput problem and runs beside k_analysis, where what a stream costs is its
    // register-time -> ONE warp per stream (no partner warps idling at the barrier through the BitAllocs), four streams per CTA;
    // fewer streams (the shards of a multi-GPU run): the per-stream chain matters more and more -> 2, 4, 8 warps, warp 0 carrying
    // the serial state and all of them sharing the line phase (scan.cuh).  Results do not depend on the choice
    // (test_images_independent_of_batching_tiling_and_smem_history runs every variant).
    int warps = a.S >= 3072 ? 1 : (a.S >= 1024 ? 2 : (a.S >= 384 ? 4 : 8));      // measured: 4096 streams 1651 (1 warp) vs 1677 ms (2), 2048: 861 vs 854, 1024: 431 (2) vs 435 (4), 512: 223 (4) vs 225 (8)
    if (const char *we = getenv("PAC_SCAN_WARPS")) { const int v = atoi(we); if (v == 1 || v == 2 || v == 4 || v == 8) warps = v; }
    {
        KTimer kt(ctx, PAC_K_SCAN);
        if (warps == 1) k_scan<T, 1><<<(a.S + kScanSoloStreams - 1) / kScanSoloStreams, kScanSoloStreams * 32, 0, LS(ctx)>>>(a);
        else if (warps == 2) k_scan<T, 2><<<a.S, 64, 0, LS(ctx)>>>(a);
        else if (warps == 4) k_scan<T, 4><<<a.S, 128, 0, LS(ctx)>>>(a);
        else k_scan<T, 8><<<a.S, 256, 0, LS(ctx)>>>(a);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return PAC_OK;
}

template <typename T>
static int launch_pack(PacCtx *ctx, PackArgs<T> &a) {
    a.ec = ctx->ec; a.bands = ctx->bands; a.M = ctx->M; a.codeLen = ctx->codeLen;
    {
        DevTables<T> tb;
        int rc = get_tables<T>(ctx, ctx->N, &tb);
        if (rc) return rc;
        a.band_of_line = tb.band_of_line;
    }
    int64_t nchunks = (int64_t)a.S * a.nb * 2;
    int64_t grid = (nchunks + kPackWarps - 1) / kPackWarps;
    int64_t maxg = (int64_t)ctx->numSMs * 8;
    if (grid > maxg) grid = maxg;
    if (grid < 1) grid = 1;
    { KTimer kt(ctx, PAC_K_PACK); k_pack<T><<<(unsigned)grid, kPackWarps * 32, 0, LS(ctx)>>>(a); }
    ctx->launches++;
    CK(cudaGetLastError());
    return PAC_OK;
}

// ------------------------------------------------------------------ encode (whole streams)
// Schedule.  Streams are cut into groups (host buffers: ONE when the batch fits a single staging buffer, else as few as the staging budget
// allows -- pageable output: two from 1024 streams on; device buffers: one),
// a group into time tiles of TB blocks.  Three internal CUDA streams:
//   sA (low priority)   k_analysis of every tile, in order
//   sB (high priority)  k_scan + k_pack of every tile, in order, one tile behind sA (double-buffered intermediates)
//   sC                  H2D of the next group's PCM / D2H of the previous group's images (host buffers only)
// The host enqueues group g+1 BEFORE it waits for group g's byte counts, so the device never drains between groups; only
// the very last tile's scan+pack is not hidden under an analysis kernel.  The caller's stream is fenced on both sides.
template <typename T>
static int encode_batch_t(PacCtx *ctx, const int16_t *pcm, int64_t stride, const int64_t *nSamples, int S, uint8_t *out,
                          int64_t cap, int64_t *outBytes, int64_t *finalState, const PacTrace *trace) {
    const int M = ctx->M, NB = ctx->bands.nBands, hdrB = header_bytes(ctx);
    const auto hostT0 = std::chrono::steady_clock::now();           // PAC_TIMELINE: host-side phases of the call
    auto hostMs = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - hostT0).count(); };
    double hostMark[5] = {0, 0, 0, 0, 0}, setupMark[4] = {0, 0, 0, 0};
    const bool pcmDev = is_device_ptr(pcm);
    // out: device memory, or PINNED host memory (k_pack then writes every chunk straight into the caller's buffer while the next
    // tile is analysed: no staging image, no copy-back phase at the end of a group, and only the outBytes[s] bytes of an image ever
    // cross PCIe), else pageable host memory (staged group by group, below).  PAC_NO_MAPPED_OUT=1 forces the staged path.
    uint8_t *outD = is_device_ptr(out) ? out : (getenv("PAC_NO_MAPPED_OUT") ? nullptr : reinterpret_cast<uint8_t *>(mapped_host_alias(out)));
    const bool outDev = outD != nullptr;
    int64_t maxBlocksAll = 0;
    for (int s = 0; s < S; s++) {
        if (nSamples[s] < 0 || nSamples[s] > stride) FAIL(PAC_E_ARG, "nSamples[%d]=%lld outside [0, strideSamples]", s, (long long)nSamples[s]);
        int64_t nb = pac_num_blocks(ctx, nSamples[s]);
        if (nb > maxBlocksAll) maxBlocksAll = nb;
    }
    if (cap < hdrB) FAIL(PAC_E_OVERFLOW, "cap smaller than the file header");
    const int64_t workBudget = 1 << 18;                       // (stream, block) items per tile buffer
    int Sg = S;
    const bool staged = !pcmDev || !outDev;
    if (staged) {
        // Host buffers are staged group by group (double-buffered when there are several).  Groups are as LARGE as the staging budget
        // allows: a group runs at the rate of its own stream count (the serial reservoir chain of k_scan is only hidden while the
        // group's own analysis lasts longer; eight groups of 256 streams measured 76 k audio-s/s where the same streams device-resident
        // ran at 101 k, four groups of 1024 streams 4 x 410 ms against 1520 ms for 4096 at once).
        size_t freeB = 0, totalB = 0;
        int64_t stagingLimit = (int64_t)64 << 30;
        if (const char *sl = getenv("PAC_STAGE_LIMIT_GB")) { const int v = atoi(sl); if (v > 0) stagingLimit = (int64_t)v << 30; }   // experiments
        // steady state (the staging buffers of an earlier call already hold the whole batch): no need to ask the driver
        const bool heldFits = (pcmDev || ctx->w_pcm.cap >= (size_t)S * stride * 4 + 16) && (outDev || ctx->w_out.cap >= (size_t)S * cap);
        if (heldFits) stagingLimit = std::max<int64_t>(stagingLimit, (int64_t)S * ((pcmDev ? 0 : stride * 4) + (outDev ? 0 : cap)));
        else if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
            const int64_t held = (int64_t)(ctx->w_pcm.cap + ctx->w_pcm2.cap + ctx->w_out.cap + ctx->w_out2.cap);   // reused, not extra
            const int64_t avail = ((int64_t)freeB + held) / 2;       // the other half stays for the tile intermediates and the caller
            if (avail < stagingLimit) stagingLimit = avail;
        }
        const int64_t per = std::max<int64_t>((pcmDev ? 0 : stride * 4) + (outDev ? 0 : cap), 1);
        int nG;
        if ((int64_t)S * per <= stagingLimit && (outDev || S < 1024)) {
            // The whole batch fits ONE staging buffer: a single group.  Nothing is double-buffered, the slab-wise H2D of the batch runs
            // ahead of the kernels, and the kernels see all streams at once -- a group of 1024 streams runs at the 1024-stream rate
            // (measured: 4 x 410 ms for 4096 x 60 s in four groups against 1520 ms for the same streams device-resident).
            nG = 1;
        } else {
            int64_t lim = stagingLimit / (2 * per);
            if (lim < 1) lim = 1;
            nG = (int)((S + lim - 1) / lim);
            if (nG < 2 && S >= 1024) nG = 2;      // pageable output: group g's images travel back under group g+1's kernels
        }
        if (const char *ge = getenv("PAC_STAGE_GROUPS")) { const int f = atoi(ge); if (f > nG) nG = f < S ? f : S; }   // tests: force staging groups
        if (trace) nG = 1;
        Sg = (S + nG - 1) / nG;
        if (Sg < 1) Sg = 1;
    }
    if (Sg > 8192) Sg = 8192;
    const int nGroups = (S + Sg - 1) / Sg;
    setupMark[0] = hostMs();
    cudaStream_t sA = ctx->sA, sB = ctx->sB, sC = ctx->sC, sM = ctx->sM, sD = ctx->sD;

    // ---- per-stream inputs and state for ALL streams, uploaded once
    CK(ctx->w_ns.ensure((size_t)S * 8));
    CK(ctx->w_state.ensure((size_t)S * sizeof(StreamState)));
    CK(ctx->w_hdr.ensure((size_t)S * hdrB));
    CK(ctx->w_ovf.ensure((size_t)S * 4));
    CK(ctx->hostState.ensure((size_t)S * (sizeof(StreamState) + 4)));         // pinned: results come back without blocking the host
    StreamState *hSt = ctx->hostState.template as<StreamState>();
    int *hOvf = reinterpret_cast<int *>(hSt + S);
    {
        std::vector<StreamState> st(S);
        for (int s = 0; s < S; s++) { st[s].extraBits = 0; st[s].bitDeposit = 0; st[s].outOffset = hdrB; st[s].reserved = 0; }   // pacfile.py:269, Huffman.py:262
        std::vector<uint8_t> hdr((size_t)S * hdrB);
        for (int s = 0; s < S; s++) build_header(ctx, nSamples[s], hdr.data() + (size_t)s * hdrB);
        CK(cudaMemcpyAsync(ctx->w_ns.p, nSamples, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->w_state.p, st.data(), (size_t)S * sizeof(StreamState), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->w_hdr.p, hdr.data(), hdr.size(), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ctx->w_ovf.p, 0, (size_t)S * 4, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));               // the pageable vectors above go out of scope
    }
    setupMark[1] = hostMs();
    CK(cudaEventRecord(ctx->evStart, ctx->stream));
    CK(cudaStreamWaitEvent(sA, ctx->evStart, 0));
    CK(cudaStreamWaitEvent(sB, ctx->evStart, 0));
    CK(cudaStreamWaitEvent(sC, ctx->evStart, 0));
    CK(cudaStreamWaitEvent(sM, ctx->evStart, 0));
    CK(cudaStreamWaitEvent(sD, ctx->evStart, 0));

    // ---- tile geometry per group, and the global tile numbering (buffer parity and events run across groups)
    struct Group { int s0, Sc, TB, nTiles, tile0; int64_t maxBlocks; std::vector<int> tstart; };
    std::vector<Group> groups(nGroups);
    int64_t nworkMax = 0;
    int totalTiles = 0;
    for (int g = 0; g < nGroups; g++) {
        Group &G = groups[g];
        G.s0 = g * Sg; G.Sc = (S - G.s0 < Sg) ? S - G.s0 : Sg;
        G.maxBlocks = 0;
        for (int s = 0; s < G.Sc; s++) { int64_t nb = pac_num_blocks(ctx, nSamples[G.s0 + s]); if (nb > G.maxBlocks) G.maxBlocks = nb; }
        int TB = (int)(workBudget / G.Sc);
        {   // at least ~16 tiles: only the LAST tile's scan+pack is not hidden under the next tile's analysis
            const int tbCap = (int)((G.maxBlocks + 15) / 16);
            if (TB > tbCap) TB = tbCap;
        }
        if (const char *tbe = getenv("PAC_TILE_BLOCKS")) TB = atoi(tbe);          // tests: force a tiling
        if (TB < 8) TB = 8;
        if (TB > G.maxBlocks) TB = (int)G.maxBlocks;
        if (trace) TB = (int)G.maxBlocks;                    // taps are copied out once per group
        // Tile boundaries.  Uniform tiles of TB blocks, except that the call's FIRST tile and LAST tile are cut into 1/8, 1/8, 1/4, 1/2:
        // the first kernel then waits for 1/128 instead of 1/16 of a group's PCM to arrive from the host, and the scan+pack of
        // the last tile -- the only one not hidden under a following analysis -- is 1/8 as long.  (Results do not depend on the tiling:
        // test_images_independent_of_batching_tiling_and_smem_history.)
        G.TB = TB;
        G.tstart.clear();
        {
            const bool uniform = trace || getenv("PAC_TILE_BLOCKS") != nullptr || TB < 64;
            const int nFull = (int)((G.maxBlocks + TB - 1) / TB);
            for (int t = 0; t < nFull; t++) {
                const int b0 = t * TB;
                const int b1 = (int)std::min<int64_t>((int64_t)b0 + TB, G.maxBlocks);
                const bool head = !uniform && g == 0 && t == 0 && nFull > 1, tail = !uniform && g == nGroups - 1 && t == nFull - 1 && nFull > 1;
                const int len = b1 - b0;
                if (head && len >= 64) { G.tstart.push_back(b0); G.tstart.push_back(b0 + len / 8); G.tstart.push_back(b0 + len / 4); G.tstart.push_back(b0 + len / 2); }
                else if (tail && len >= 64) { G.tstart.push_back(b0); G.tstart.push_back(b0 + len / 2); G.tstart.push_back(b0 + len / 2 + len / 4); G.tstart.push_back(b0 + len / 2 + len / 4 + len / 8); }
                else G.tstart.push_back(b0);
            }
            G.tstart.push_back((int)G.maxBlocks);
        }
        G.nTiles = (int)G.tstart.size() - 1; G.tile0 = totalTiles;
        totalTiles += G.nTiles;
        if ((int64_t)G.Sc * TB > nworkMax) nworkMax = (int64_t)G.Sc * TB;
    }
    // PAC_MDCT_AHEAD=1 (fp32 mode, experiment kept as an option): k_mdct_enc of tiles t+1, t+2 runs on its own stream BESIDE the analysis
    // of tile t instead of in front of the analysis of its own tile (one more set of tile buffers).  Measured: 4096 x 60 s 1620.5 ->
    // 1619.0 ms, 512 x 60 s 217.9 -> 214.3 ms -- the SMs are busy either way, a step costs the sum of its kernels' work; off by
    // default because the per-kernel times of the overlapped MDCT then say nothing (bench.py's stage split reads them).
    const bool ahead = sizeof(T) == 4 && !trace && getenv("PAC_MDCT_AHEAD") && atoi(getenv("PAC_MDCT_AHEAD")) != 0 && totalTiles > 2;
    const int NBUF = totalTiles > 1 ? (ahead ? 3 : 2) : 1;      // (a third set of tile buffers without the MDCT running ahead measured no gain: 512 x 60 s 204.8 vs 205.1 ms)
    setupMark[2] = hostMs();
    // Pinned host output of a single-group batch: k_pack writes into a device image and k_drain (pack.cuh) moves every tile's finished
    // byte ranges to the caller's image beside the next tile's analysis.  PAC_NO_DRAIN=1: k_pack writes across PCIe itself, as it does
    // for multi-group batches.
    const bool drain = outDev && !is_device_ptr(out) && nGroups == 1 && !trace && totalTiles > 1 && !getenv("PAC_NO_DRAIN");
    uint8_t *drainImg = nullptr;
    long long *drainOff = nullptr;
    if (drain) {
        CK(ctx->w_out.ensure((size_t)S * cap + 32));
        CK(ctx->w_toff.ensure((size_t)totalTiles * S * 2 * sizeof(long long)));
        const uintptr_t base = reinterpret_cast<uintptr_t>(ctx->w_out.p);
        drainImg = ctx->w_out.template as<uint8_t>() + ((reinterpret_cast<uintptr_t>(outD) & 15) + 16 - (base & 15)) % 16;   // same alignment mod 16 as the host image
        drainOff = ctx->w_toff.template as<long long>();
    }
    const size_t szLines = (size_t)nworkMax * 2 * M * sizeof(T), szBand = (size_t)nworkMax * 2 * kMaxBands * sizeof(T);
    CK(ctx->w_lines.ensure(szLines * NBUF));
    CK(ctx->w_smr.ensure(szBand * NBUF));
    CK(ctx->w_bmax.ensure(szBand * NBUF));
    CK(ctx->w_osc.ensure((size_t)nworkMax * 2 * NBUF));
    CK(ctx->w_lrms.ensure((size_t)nworkMax * 4 * NBUF));
    CK(ctx->w_ba.ensure((size_t)nworkMax * 2 * kMaxBands));
    CK(ctx->w_sf.ensure((size_t)nworkMax * 2 * kMaxBands));
    CK(ctx->w_tid.ensure((size_t)nworkMax * 2));
    CK(ctx->w_nby.ensure((size_t)nworkMax * 2 * 4));
    CK(ctx->w_coff.ensure((size_t)nworkMax * 2 * 8));
    if (trace) { CK(ctx->w_trE.ensure((size_t)nworkMax * 8)); CK(ctx->w_trD.ensure((size_t)nworkMax * 8)); }
    // NB: evA is created WITH timing on purpose.  Measured on B200 (512 x 60 s, device-resident): 354 ms per call, against
    // 425 ms when evA carries cudaEventDisableTiming -- the cross-stream wait of sB on a timestamped record lets the
    // high-priority scan kernel in promptly; with the light-weight event it trails the next analysis kernel.
    static const int evaTiming = getenv("PAC_EVA_TIMING") ? atoi(getenv("PAC_EVA_TIMING")) : 1;
    while ((int)ctx->evA.size() < totalTiles) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, (evaTiming & 1) ? cudaEventDefault : cudaEventDisableTiming)); ctx->evA.push_back(e); }
    while ((int)ctx->evB.size() < totalTiles) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, (evaTiming & 2) ? cudaEventDefault : cudaEventDisableTiming)); ctx->evB.push_back(e); }
    while ((int)ctx->evM.size() < totalTiles) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->evM.push_back(e); }
    while ((int)ctx->evG.size() < nGroups) { cudaEvent_t e; CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->evG.push_back(e); }
    if (!pcmDev) { CK(ctx->w_pcm.ensure((size_t)Sg * stride * 4 + 16)); if (nGroups > 1) CK(ctx->w_pcm2.ensure((size_t)Sg * stride * 4 + 16)); }
    if (!outDev) { CK(ctx->w_out.ensure((size_t)Sg * cap)); if (nGroups > 1) CK(ctx->w_out2.ensure((size_t)Sg * cap)); }

    setupMark[3] = hostMs();
    // PAC_TIMELINE=1: print when each group's copies and kernels ran (ms since the start of the call)
    struct Mark { cudaEvent_t e; const char *what; int g; };
    std::vector<Mark> marks;
    const bool timeline = getenv("PAC_TIMELINE") != nullptr;
    cudaEvent_t tl0 = nullptr;
    auto mark = [&](cudaStream_t st, const char *what, int g) {
        if (!timeline) return;
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); marks.push_back({e, what, g});
    };
    if (timeline) { cudaEventCreate(&tl0); cudaEventRecord(tl0, ctx->stream); }
    hostMark[0] = hostMs();                                         // set-up done (state upload, buffers, events)

    // H2D of a group's PCM goes in time slabs (one per tile, all streams of the group: a strided 2-D copy), each with its own
    // event, so that the analysis of tile t starts as soon as samples < (t+1)*TB*M have landed -- the first kernel of the
    // call waits for 1/nTiles of a group, not for a whole group.
    auto h2d_group = [&](int g) -> cudaError_t {
        const Group &G = groups[g];
        DBuf &buf = (g & 1) ? ctx->w_pcm2 : ctx->w_pcm;
        cudaError_t e = cudaSuccess;
        if (g >= 2) e = cudaStreamWaitEvent(sC, ctx->evA[groups[g - 2].tile0 + groups[g - 2].nTiles - 1], 0);   // the buffer's previous reader is done
        if (e != cudaSuccess) return e;
        std::vector<cudaEvent_t> &evs = ctx->evS[g & 1];
        while ((int)evs.size() < G.nTiles) { cudaEvent_t ev; e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming); if (e != cudaSuccess) return e; evs.push_back(ev); }
        mark(sC, "h2d begin", g);
        for (int t = 0; t < G.nTiles; t++) {
            const int64_t a0 = (int64_t)G.tstart[t] * M;
            int64_t a1 = (t == G.nTiles - 1) ? stride : (int64_t)G.tstart[t + 1] * M;
            if (a1 > stride) a1 = stride;
            if (a1 > a0) {
                e = cudaMemcpy2DAsync(buf.template as<char>() + a0 * 4, (size_t)stride * 4, reinterpret_cast<const char *>(pcm + (int64_t)G.s0 * stride * 2) + a0 * 4,
                                      (size_t)stride * 4, (size_t)(a1 - a0) * 4, (size_t)G.Sc, cudaMemcpyHostToDevice, sC);
                if (e != cudaSuccess) return e;
            }
            e = cudaEventRecord(evs[t], sC);
            if (e != cudaSuccess) return e;
        }
        mark(sC, "h2d end", g);
        return cudaSuccess;
    };

    int status = PAC_OK;
    auto enqueue_group = [&](int g) -> int {
        const Group &G = groups[g];
        const int Sc = G.Sc, s0 = G.s0, TB = G.TB;
        const int64_t nwork = (int64_t)Sc * TB;
        const int16_t *d_pcm;
        if (pcmDev) d_pcm = pcm + (int64_t)s0 * stride * 2;
        else {
            d_pcm = ((g & 1) ? ctx->w_pcm2 : ctx->w_pcm).template as<int16_t>();
            if (g + 1 < nGroups) CK(h2d_group(g + 1));                                     // next group's copy overlaps this group's kernels
        }
        uint8_t *d_out;
        if (outDev) d_out = outD + (int64_t)s0 * cap;
        else {
            d_out = ((g & 1) ? ctx->w_out2 : ctx->w_out).template as<uint8_t>();
            if (g >= 2) CK(cudaStreamWaitEvent(sB, ctx->evD[g & 1], 0));                  // its previous contents have been copied out
        }
        for (int t = 0; t < G.nTiles; t++) {
            const int b0 = G.tstart[t];
            const int tile = G.tile0 + t;
            const int nb = G.tstart[t + 1] - b0;
            const int pbuf = tile % NBUF;
            AnalysisArgs<T> aa{};
            aa.pcm = d_pcm; aa.strideSamples = stride; aa.nSamples = ctx->w_ns.template as<int64_t>() + s0; aa.blocks = nullptr;
            aa.S = Sc; aa.b0 = b0; aa.nb = nb; aa.nwork = (int64_t)Sc * nb;
            aa.lines = reinterpret_cast<T *>(ctx->w_lines.template as<char>() + szLines * pbuf);
            aa.smr = reinterpret_cast<T *>(ctx->w_smr.template as<char>() + szBand * pbuf);
            aa.bmax = reinterpret_cast<T *>(ctx->w_bmax.template as<char>() + szBand * pbuf);
            aa.oscale = ctx->w_osc.template as<uint8_t>() + (size_t)nworkMax * 2 * pbuf;
            aa.lrms = ctx->w_lrms.template as<uint32_t>() + (size_t)nworkMax * pbuf;
            aa.dbg_mdct = nullptr; aa.dbg_bthr = nullptr;
            cudaStream_t sFirst = ahead ? sM : sA;                                        // the stream of the tile's first kernel
            if (tile >= NBUF) CK(cudaStreamWaitEvent(sFirst, ctx->evB[tile - NBUF], 0));   // tile buffer reuse
            if (!pcmDev) CK(cudaStreamWaitEvent(sFirst, ctx->evS[g & 1][t], 0));          // this tile's samples have landed
            int rc;
            if (ahead) {
                ctx->launchStream = sM;
                rc = launch_mdct(ctx, aa);
                if (rc) { ctx->launchStream = nullptr; return rc; }
                CK(cudaEventRecord(ctx->evM[tile], sM));
                CK(cudaStreamWaitEvent(sA, ctx->evM[tile], 0));
            }
            if (t == 0) mark(sA, "analysis begin", g);
            ctx->launchStream = sA;
            rc = launch_analysis<T>(ctx, aa, !ahead);
            if (rc) { ctx->launchStream = nullptr; return rc; }
            CK(cudaEventRecord(ctx->evA[tile], sA));
            if (t == G.nTiles - 1) mark(sA, "analysis end", g);
            CK(cudaStreamWaitEvent(sB, ctx->evA[tile], 0));
            ctx->launchStream = sB;
            ScanArgs<T> sa{};
            sa.S = Sc; sa.b0 = b0; sa.nb = nb; sa.nSamples = aa.nSamples; sa.state = ctx->w_state.template as<StreamState>() + s0;
            sa.lines = aa.lines; sa.smr = aa.smr; sa.bmax = aa.bmax; sa.lrms = aa.lrms;
            sa.ba = ctx->w_ba.template as<uint8_t>(); sa.sf = ctx->w_sf.template as<uint8_t>(); sa.tableID = ctx->w_tid.template as<uint8_t>();
            sa.nbytes = ctx->w_nby.template as<uint32_t>(); sa.chunkOff = ctx->w_coff.template as<long long>();
            sa.trExtra = trace ? ctx->w_trE.template as<long long>() : nullptr; sa.trDeposit = trace ? ctx->w_trD.template as<long long>() : nullptr;
            if (drain) { sa.tileBeg = drainOff + (size_t)tile * S * 2; sa.tileEnd = sa.tileBeg + S; }
            if ((rc = launch_scan<T>(ctx, sa))) { ctx->launchStream = nullptr; return rc; }
            PackArgs<T> pa{};
            pa.S = Sc; pa.b0 = b0; pa.nb = nb; pa.nSamples = aa.nSamples;
            pa.lines = aa.lines; pa.ba = sa.ba; pa.sf = sa.sf; pa.tableID = sa.tableID; pa.oscale = aa.oscale; pa.lrms = aa.lrms;
            pa.nbytes = sa.nbytes; pa.chunkOff = sa.chunkOff; pa.out = drain ? drainImg : d_out; pa.cap = cap; pa.perChunk = 0;
            pa.overflow = ctx->w_ovf.template as<int>() + s0; pa.o_mant = nullptr;
            if (trace && trace->mant) {
                CK(ctx->w_misc.ensure((size_t)nwork * 2 * M * 4));
                CK(cudaMemsetAsync(ctx->w_misc.p, 0, (size_t)nwork * 2 * M * 4, sB));
                pa.o_mant = ctx->w_misc.template as<int32_t>();
            }
            pa.header = ctx->w_hdr.template as<uint8_t>() + (size_t)s0 * hdrB; pa.headerBytes = hdrB;
            rc = launch_pack<T>(ctx, pa);
            ctx->launchStream = nullptr;
            if (rc) return rc;
            CK(cudaEventRecord(ctx->evB[tile], sB));
            if (drain) {
                CK(cudaStreamWaitEvent(sD, ctx->evB[tile], 0));
                DrainArgs da{};
                da.img = drainImg; da.dst = d_out; da.cap = cap; da.tileBeg = sa.tileBeg; da.tileEnd = sa.tileEnd; da.S = Sc;
                int dgrid = ctx->numSMs / 4;
                if (dgrid > (Sc + kDrainWarps - 1) / kDrainWarps) dgrid = (Sc + kDrainWarps - 1) / kDrainWarps;
                if (dgrid < 1) dgrid = 1;
                k_drain<<<dgrid, kDrainWarps * 32, 0, sD>>>(da);
                ctx->launches++;
                CK(cudaGetLastError());
            }
            if (trace) {      // single tile (TB == maxBlocks): copy the taps of this stream group
                CK(cudaStreamSynchronize(sB));
                const int64_t B = maxBlocksAll;
                std::vector<T> hl, hs;
                std::vector<uint8_t> hba((size_t)nwork * 2 * kMaxBands), hsf((size_t)nwork * 2 * kMaxBands), htid((size_t)nwork * 2), hosc((size_t)nwork * 2);
                std::vector<uint32_t> hlr(nwork), hnby((size_t)nwork * 2);
                std::vector<long long> hE(nwork), hD(nwork);
                std::vector<int32_t> hm;
                if (trace->mant) { hm.resize((size_t)nwork * 2 * M); CK(cudaMemcpy(hm.data(), pa.o_mant, hm.size() * 4, cudaMemcpyDeviceToHost)); }
                if (trace->lines) { hl.resize((size_t)nwork * 2 * M); CK(cudaMemcpy(hl.data(), aa.lines, hl.size() * sizeof(T), cudaMemcpyDeviceToHost)); }
                if (trace->smr) { hs.resize((size_t)nwork * 2 * kMaxBands); CK(cudaMemcpy(hs.data(), aa.smr, hs.size() * sizeof(T), cudaMemcpyDeviceToHost)); }
                CK(cudaMemcpy(hba.data(), sa.ba, hba.size(), cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(hsf.data(), sa.sf, hsf.size(), cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(htid.data(), sa.tableID, htid.size(), cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(hosc.data(), aa.oscale, hosc.size(), cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(hlr.data(), aa.lrms, hlr.size() * 4, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(hnby.data(), sa.nbytes, hnby.size() * 4, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(hE.data(), sa.trExtra, hE.size() * 8, cudaMemcpyDeviceToHost));
                CK(cudaMemcpy(hD.data(), sa.trDeposit, hD.size() * 8, cudaMemcpyDeviceToHost));
                for (int s = 0; s < Sc; s++) {
                    int64_t nbs = pac_num_blocks(ctx, nSamples[s0 + s]);
                    for (int64_t b = 0; b < nbs; b++) {
                        int64_t w = (int64_t)s * nb + b, gi = (int64_t)(s0 + s) * B + b;
                        if (trace->lrms) trace->lrms[gi] = (int32_t)hlr[w];
                        if (trace->extraBits) trace->extraBits[gi] = hE[w];
                        if (trace->bitDeposit) trace->bitDeposit[gi] = hD[w];
                        for (int ch = 0; ch < 2; ch++) {
                            if (trace->oscale) trace->oscale[gi * 2 + ch] = hosc[w * 2 + ch];
                            if (trace->tableID) trace->tableID[gi * 2 + ch] = htid[w * 2 + ch];
                            if (trace->nbytes) trace->nbytes[gi * 2 + ch] = (int32_t)hnby[w * 2 + ch];
                            for (int bd = 0; bd < NB; bd++) {
                                if (trace->ba) trace->ba[(gi * 2 + ch) * NB + bd] = hba[(w * 2 + ch) * kMaxBands + bd];
                                if (trace->sf) trace->sf[(gi * 2 + ch) * NB + bd] = hsf[(w * 2 + ch) * kMaxBands + bd];
                                if (trace->smr) trace->smr[(gi * 2 + ch) * NB + bd] = (double)hs[(w * 2 + ch) * kMaxBands + bd];
                            }
                            if (trace->lines)
                                for (int i = 0; i < M; i++) trace->lines[(gi * 2 + ch) * M + i] = (double)hl[(w * 2 + ch) * M + i];
                            if (trace->mant) memcpy(trace->mant + (gi * 2 + ch) * M, hm.data() + (w * 2 + ch) * M, (size_t)M * 4);
                        }
                    }
                }
            }
        }
        // this group's final per-stream state -> pinned host memory, behind its last pack
        CK(cudaMemcpyAsync(hSt + s0, ctx->w_state.template as<StreamState>() + s0, (size_t)Sc * sizeof(StreamState), cudaMemcpyDeviceToHost, sB));
        CK(cudaMemcpyAsync(hOvf + s0, ctx->w_ovf.template as<int>() + s0, (size_t)Sc * 4, cudaMemcpyDeviceToHost, sB));
        CK(cudaEventRecord(ctx->evG[g], sB));
        mark(sB, "pack end", g);
        return PAC_OK;
    };
    auto collect_group = [&](int g) -> int {
        const Group &G = groups[g];
        CK(cudaEventSynchronize(ctx->evG[g]));
        int64_t maxBytes = 0;
        for (int s = G.s0; s < G.s0 + G.Sc; s++) {
            outBytes[s] = hSt[s].outOffset;
            if (finalState) { finalState[s * 2] = hSt[s].bitDeposit; finalState[s * 2 + 1] = hSt[s].extraBits; }
            if (hOvf[s] || hSt[s].outOffset > cap) { status = PAC_E_OVERFLOW; outBytes[s] = -hSt[s].outOffset; }
            else if (hSt[s].outOffset > maxBytes) maxBytes = hSt[s].outOffset;
        }
        if (!outDev) {                                    // the kernels of this group are complete (event synchronised above)
            const uint8_t *d_out = ((g & 1) ? ctx->w_out2 : ctx->w_out).template as<uint8_t>();
            mark(sC, "d2h begin", g);
            if (maxBytes > 0)
                CK(cudaMemcpy2DAsync(out + (int64_t)G.s0 * cap, (size_t)cap, d_out, (size_t)cap, (size_t)maxBytes, (size_t)G.Sc, cudaMemcpyDeviceToHost, sC));
            mark(sC, "d2h end", g);
            CK(cudaEventRecord(ctx->evD[g & 1], sC));
        }
        return PAC_OK;
    };

    if (!pcmDev) CK(h2d_group(0));
    hostMark[1] = hostMs();                                         // first group's copies enqueued
    int rc = PAC_OK;
    for (int g = 0; g < nGroups && rc == PAC_OK; g++) {
        rc = enqueue_group(g);
        if (g == 0) hostMark[2] = hostMs();                         // first group's kernels enqueued
        if (rc == PAC_OK && g >= 1) rc = collect_group(g - 1);
    }
    if (rc == PAC_OK) rc = collect_group(nGroups - 1);
    hostMark[3] = hostMs();                                         // last group's byte counts are on the host
    // fence: the caller's stream continues after everything issued here (also on the error path, so buffers can be reused)
    cudaStreamSynchronize(sM); cudaStreamSynchronize(sA); cudaStreamSynchronize(sB); cudaStreamSynchronize(sC); cudaStreamSynchronize(sD);
    hostMark[4] = hostMs();
    if (timeline) {
        fprintf(stderr, "[pac timeline] host set-up: grouping %.2f ms, state upload %.2f, tile geometry %.2f, buffers+events %.2f\n", setupMark[0], setupMark[1], setupMark[2], setupMark[3]);
        fprintf(stderr, "[pac timeline] host: set-up %.2f ms, first copies enqueued %.2f, first group's kernels enqueued %.2f, results collected %.2f, streams drained %.2f (%d group(s))\n",
                hostMark[0], hostMark[1], hostMark[2], hostMark[3], hostMark[4], nGroups);
        for (const Mark &m : marks) { float ms = 0; cudaEventElapsedTime(&ms, tl0, m.e); fprintf(stderr, "[pac timeline] group %d %-15s %9.2f ms\n", m.g, m.what, ms); cudaEventDestroy(m.e); }
        cudaEventDestroy(tl0);
    }
    if (rc != PAC_OK) return rc;
    if (status == PAC_E_OVERFLOW) ctx->err = "output capacity too small for at least one stream (outBytes[s] = -needed)";
    return status;
}

extern "C" int pac_encode_batch(PacCtx *ctx, const int16_t *pcm, int64_t strideSamples, const int64_t *nSamples, int S,
                                uint8_t *out, int64_t cap, int64_t *outBytes, int64_t *finalState, const PacTrace *trace) {
    if (!ctx) return PAC_E_ARG;
    if (!pcm || !nSamples || !out || !outBytes || S <= 0 || strideSamples < 0) FAIL(PAC_E_ARG, "bad arguments to pac_encode_batch");
    CK(cudaSetDevice(ctx->device));
    if (ctx->precision == PAC_PRECISION_FP64) return encode_batch_t<double>(ctx, pcm, strideSamples, nSamples, S, out, cap, outBytes, finalState, trace);
    return encode_batch_t<float>(ctx, pcm, strideSamples, nSamples, S, out, cap, outBytes, finalState, trace);
}

// ------------------------------------------------------------------ decode (whole streams)
extern "C" int64_t pac_decode_bound(PacCtx *ctx, int64_t nbytes) {
    if (!ctx || nbytes < 0) return PAC_E_ARG;
    int64_t minBlock = 2 * (4 + (ctx->ec.fixedBits + 7) / 8);
    return (nbytes / minBlock + 2) * ctx->M;
}

template <typename T, int LOGM>
static int launch_synth_t(PacCtx *ctx, SynthArgs<T> &a, int64_t grid) {
    size_t smem = sizeof(SynthSmem<T, LOGM>);
    CK(cudaFuncSetAttribute(k_synth<T, LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { KTimer kt(ctx, PAC_K_SYNTH); k_synth<T, LOGM><<<(unsigned)grid, (1 << LOGM) / 4, smem, ctx->stream>>>(a); }
    ctx->launches++;
    CK(cudaGetLastError());
    return PAC_OK;
}

template <typename T>
static int launch_synth(PacCtx *ctx, SynthArgs<T> &a, int64_t grid) {
    a.bands = ctx->bands;
    int rc = get_tables<T>(ctx, ctx->N, &a.tab);
    if (rc) return rc;
    if (ctx->LOGM == 10) return launch_synth_t<T, 10>(ctx, a, grid);
    if (ctx->LOGM == 9) return launch_synth_t<T, 9>(ctx, a, grid);
    FAIL(PAC_E_ARG, "unsupported nMDCTLines");
}

constexpr int kRetryCounting = -101;      // internal: the header's block count was too small, walk the chains first

template <typename T>
static int decode_batch_t(PacCtx *ctx, const uint8_t *pac, const int64_t *pacBeg, const int64_t *pacLen, int S, int16_t *pcm, int64_t stride,
                          int64_t *nSamplesOut, int64_t *hdrNumSamples, int32_t *hdrSampleRate, bool countFirst) {
    // countFirst = false: the number of blocks of a stream is taken from its header's sample count (what this encoder and the
    // reference write, pacfile.py:237-246) and only VERIFIED by the one walk that also fills the chunk index; a stream whose chain
    // holds more blocks than its header promises makes the call return kRetryCounting, and the caller repeats it with
    // countFirst = true (a first walk counts the blocks of every chain, as round 1 always did).
    const int M = ctx->M, hdrB = header_bytes(ctx);
    const bool pacDev = is_device_ptr(pac), pcmDev = is_device_ptr(pcm);
    const int64_t minBlock = 2 * (4 + (ctx->ec.fixedBits + 7) / 8);
    for (int s = 0; s < S; s++)
        if (pacLen[s] < hdrB) FAIL(PAC_E_FORMAT, "stream %d shorter than a PAC header", s);
    // ---- input: device images are used in place; host images are packed into one staging buffer
    const uint8_t *d_pac = pac;
    std::vector<int64_t> beg(pacBeg, pacBeg + S), len(pacLen, pacLen + S);
    int64_t maxLen = 0;
    if (!pacDev) {
        int64_t tot = 0;
        for (int s = 0; s < S; s++) { beg[s] = tot; tot += (len[s] + 15) & ~(int64_t)15; }
        CK(ctx->w_misc.ensure((size_t)tot + 16));
        for (int s = 0; s < S; s++)
            CK(cudaMemcpyAsync(ctx->w_misc.as<uint8_t>() + beg[s], pac + pacBeg[s], (size_t)len[s], cudaMemcpyHostToDevice, ctx->stream));
        d_pac = ctx->w_misc.as<uint8_t>();
    }
    for (int s = 0; s < S; s++) if (len[s] > maxLen) maxLen = len[s];
    CK(ctx->w_ns.ensure((size_t)S * 16));
    int64_t *d_beg = ctx->w_ns.as<int64_t>(), *d_len = d_beg + S;
    CK(cudaMemcpyAsync(d_beg, beg.data(), (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d_len, len.data(), (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    // ---- headers (pacfile.py:123-151) and a first walk of every stream's chunk chain (counts blocks, finds truncation)
    std::vector<uint8_t> hdr((size_t)S * hdrB);
    std::vector<int32_t> nblk(S), stt(S);
    CK(ctx->w_hdr.ensure(hdr.size()));
    CK(ctx->w_misc2.ensure((size_t)S * 4));   // nBlocks
    CK(ctx->w_misc3.ensure((size_t)S * 4));   // status
    k_gather_headers<<<(unsigned)((hdr.size() + 255) / 256), 256, 0, ctx->stream>>>(d_pac, d_beg, S, hdrB, ctx->w_hdr.as<uint8_t>());
    ctx->launches++;
    CK(cudaGetLastError());
    IndexArgs ia{};
    ia.pac = d_pac; ia.pacBeg = d_beg; ia.pacLen = d_len; ia.S = S; ia.hdrBytes = hdrB;
    ia.maxBlocks = (int)((maxLen - hdrB) / minBlock + 1);
    ia.chunkPos = nullptr; ia.chunkLen = nullptr;
    ia.nBlocks = ctx->w_misc2.as<int32_t>(); ia.status = ctx->w_misc3.as<int32_t>();
    if (countFirst) {
        { KTimer kt(ctx, PAC_K_INDEX); k_index<<<(S + kIdxWarps - 1) / kIdxWarps, 32 * kIdxWarps, 0, ctx->stream>>>(ia); }
        ctx->launches++;
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(nblk.data(), ia.nBlocks, (size_t)S * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(stt.data(), ia.status, (size_t)S * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaMemcpyAsync(hdr.data(), ctx->w_hdr.p, hdr.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    auto rd32 = [](const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); };
    auto rd16 = [](const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); };
    int maxBlocks = 1;
    for (int s = 0; s < S; s++) {
        const uint8_t *h = hdr.data() + (size_t)s * hdrB;
        if (memcmp(h, "PAC ", 4)) FAIL(PAC_E_FORMAT, "stream %d: Tried to read a non-PAC file into a PACFile object", s);   // pacfile.py:130
        if ((int)rd16(h + 8) != 2 || (int)rd32(h + 14) != M || (int)rd16(h + 18) != ctx->p.nScaleBits ||
            (int)rd16(h + 20) != ctx->p.nMantSizeBits || (int)rd32(h + 22) != ctx->bands.nBands)
            FAIL(PAC_E_FORMAT, "stream %d: header does not match this context's coding parameters", s);
        for (int b = 0; b < ctx->bands.nBands; b++)
            if ((int)rd16(h + 26 + 2 * b) != ctx->nLines[b]) FAIL(PAC_E_FORMAT, "stream %d: band layout differs from this context's", s);
        if (hdrSampleRate) hdrSampleRate[s] = (int32_t)rd32(h + 4);
        if (hdrNumSamples) hdrNumSamples[s] = rd32(h + 10);
        if (!countFirst) {
            // blocks the header promises: n samples were coded as ceil(n/M) + 1 blocks, and the header holds n, or n + M when n is a
            // multiple of M (the padding rule as written, pacfile.py:240-242)
            const int64_t hn = rd32(h + 10);
            int64_t hint = (hn % M == 0) ? hn / M : hn / M + 2;
            const int64_t byLen = (len[s] - hdrB) / minBlock + 1;
            if (hint > byLen) hint = byLen;
            if (hint < 1) hint = 1;
            if (hint * M > stride) return kRetryCounting;      // the real count may still fit: count first
            nblk[s] = (int32_t)hint; stt[s] = 0;
        }
        if (stt[s] == kIdxBoundExceeded) FAIL(PAC_E_FORMAT, "stream %d: more chunks than its length allows", s);
        if (stt[s]) FAIL(PAC_E_FORMAT, "stream %d: Only read a partial block of coded PACFile data", s);   // pacfile.py:184
        if ((int64_t)nblk[s] * M > stride) FAIL(PAC_E_OVERFLOW, "stream %d decodes to %lld samples > strideSamples", s, (long long)nblk[s] * M);
        if (nblk[s] > maxBlocks) maxBlocks = nblk[s];
    }
    // ---- stream groups keep the mantissa-code intermediate bounded (~8 GB)
    const int64_t perStream = (int64_t)maxBlocks * 2 * M * (int64_t)sizeof(uint16_t);
    int Sg = (int)(((int64_t)8 << 30) / perStream);
    if (Sg < 1) Sg = 1;
    if (Sg > S) Sg = S;
    for (int s0 = 0; s0 < S; s0 += Sg) {
        const int Sc = (S - s0 < Sg) ? S - s0 : Sg;
        // ---- chunk index of this group (second walk, now into exactly sized arrays)
        const int64_t nblkAll = (int64_t)Sc * maxBlocks;
        CK(ctx->w_coff.ensure((size_t)nblkAll * 2 * 8));
        CK(ctx->w_nby.ensure((size_t)nblkAll * 2 * 4));
        CK(ctx->w_misc5.ensure((size_t)Sc * 8));   // nBlocks, status of the second walk
        IndexArgs ig = ia;
        ig.pacBeg = d_beg + s0; ig.pacLen = d_len + s0; ig.S = Sc; ig.maxBlocks = maxBlocks;
        ig.chunkPos = ctx->w_coff.as<int64_t>(); ig.chunkLen = ctx->w_nby.as<int32_t>();
        ig.nBlocks = ctx->w_misc5.as<int32_t>(); ig.status = ig.nBlocks + Sc;
        { KTimer kt(ctx, PAC_K_INDEX); k_index<<<(Sc + kIdxWarps - 1) / kIdxWarps, 32 * kIdxWarps, 0, ctx->stream>>>(ig); }
        ctx->launches++;
        CK(cudaGetLastError());
        // ---- unpack (dequantisation happens in the synthesis kernel, all threads in parallel)
        CK(ctx->w_lines.ensure((size_t)nblkAll * 2 * M * sizeof(uint16_t)));
        CK(ctx->w_ba.ensure((size_t)nblkAll * 2 * kMaxBands * sizeof(uint16_t)));
        CK(ctx->w_sf.ensure((size_t)nblkAll * 2));
        CK(ctx->w_lrms.ensure((size_t)nblkAll * 4));
        CK(ctx->w_misc6.ensure((size_t)Sc * 4));
        CK(cudaMemsetAsync(ctx->w_misc6.p, 0, (size_t)Sc * 4, ctx->stream));
        UnpackArgs<T> ua{};
        ua.pac = d_pac; ua.chunkPos = ig.chunkPos; ua.chunkLen = ig.chunkLen; ua.nBlocks = ig.nBlocks;
        ua.S = Sc; ua.maxBlocks = maxBlocks; ua.M = M;
        ua.nScaleBits = ctx->p.nScaleBits; ua.nMantSizeBits = ctx->p.nMantSizeBits; ua.nTableIDBits = 4;   // pacfile.py:189
        ua.codes = ctx->w_lines.as<uint16_t>(); ua.meta = ctx->w_ba.as<uint16_t>(); ua.oscale = ctx->w_sf.as<uint8_t>();
        ua.lrms = ctx->w_lrms.as<uint32_t>(); ua.err = ctx->w_misc6.as<int32_t>();
        ua.dt = ctx->dt; ua.bands = ctx->bands;
        {
            int64_t nchunk = nblkAll * 2;
            int64_t grid = (nchunk + kUnpackThreads - 1) / kUnpackThreads;
            if (grid > (int64_t)ctx->numSMs * 16) grid = (int64_t)ctx->numSMs * 16;
            { KTimer kt(ctx, PAC_K_UNPACK); k_unpack<T><<<(unsigned)grid, kUnpackThreads, 0, ctx->stream>>>(ua); }
            ctx->launches++;
            CK(cudaGetLastError());
        }
        // ---- synthesis
        int16_t *d_pcm;
        if (pcmDev) d_pcm = pcm + (int64_t)s0 * stride * 2;
        else { CK(ctx->w_pcm.ensure((size_t)Sc * stride * 4)); d_pcm = ctx->w_pcm.as<int16_t>(); }
        CK(ctx->w_misc4.ensure((size_t)Sc * 8));
        SynthArgs<T> sa{};
        sa.lines = nullptr; sa.codes = ua.codes; sa.meta = ua.meta; sa.oscale = ua.oscale; sa.largestScale = (1 << ctx->p.nScaleBits) - 1;
        sa.lrms = ua.lrms; sa.nBlocks = ig.nBlocks; sa.S = Sc; sa.maxBlocks = maxBlocks; sa.run = 16;
        sa.pcm = d_pcm; sa.strideSamples = stride; sa.nSamplesOut = ctx->w_misc4.as<int64_t>(); sa.rawOut = nullptr;
        int runsPerStream = (maxBlocks + sa.run) / sa.run;
        int rc = launch_synth<T>(ctx, sa, (int64_t)Sc * runsPerStream);
        if (rc) return rc;
        std::vector<int32_t> ib((size_t)Sc * 2);
        CK(cudaMemcpyAsync(stt.data(), ctx->w_misc6.p, (size_t)Sc * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(ib.data(), ig.nBlocks, (size_t)Sc * 8, cudaMemcpyDeviceToHost, ctx->stream));     // blocks found, walk status
        CK(cudaStreamSynchronize(ctx->stream));
        int64_t maxS = 0;
        for (int s = 0; s < Sc; s++) {
            if (ib[Sc + s] == kIdxBoundExceeded) {
                if (!countFirst) return kRetryCounting;
                FAIL(PAC_E_FORMAT, "stream %d: more chunks than its length allows", s0 + s);
            }
            if (ib[Sc + s]) FAIL(PAC_E_FORMAT, "stream %d: Only read a partial block of coded PACFile data", s0 + s);   // pacfile.py:184
            if (stt[s]) FAIL(PAC_E_FORMAT, "stream %d: malformed chunk (bad table ID or code)", s0 + s);
            nblk[s0 + s] = ib[s];
            nSamplesOut[s0 + s] = (int64_t)nblk[s0 + s] * M;
            if (nSamplesOut[s0 + s] > maxS) maxS = nSamplesOut[s0 + s];
        }
        if (!pcmDev && maxS > 0)
            CK(cudaMemcpy2D(pcm + (int64_t)s0 * stride * 2, (size_t)stride * 4, d_pcm, (size_t)stride * 4, (size_t)maxS * 4, (size_t)Sc, cudaMemcpyDeviceToHost));
    }
    return PAC_OK;
}

extern "C" int pac_decode_batch_strided(PacCtx *ctx, const uint8_t *pac, const int64_t *pacBeg, const int64_t *pacLen, int S, int16_t *pcm,
                                        int64_t strideSamples, int64_t *nSamplesOut, int64_t *hdrNumSamples, int32_t *hdrSampleRate) {
    if (!ctx) return PAC_E_ARG;
    if (!pac || !pacBeg || !pacLen || !pcm || !nSamplesOut || S <= 0) FAIL(PAC_E_ARG, "bad arguments to pac_decode_batch");
    CK(cudaSetDevice(ctx->device));
    for (int pass = 0; pass < 2; pass++) {
        const bool countFirst = pass == 1;
        const int rc = ctx->precision == PAC_PRECISION_FP64
                           ? decode_batch_t<double>(ctx, pac, pacBeg, pacLen, S, pcm, strideSamples, nSamplesOut, hdrNumSamples, hdrSampleRate, countFirst)
                           : decode_batch_t<float>(ctx, pac, pacBeg, pacLen, S, pcm, strideSamples, nSamplesOut, hdrNumSamples, hdrSampleRate, countFirst);
        if (rc != kRetryCounting) return rc;
    }
    FAIL(PAC_E_FORMAT, "chunk chain could not be indexed");
}

extern "C" int pac_decode_batch(PacCtx *ctx, const uint8_t *pac, const int64_t *pacOff, int S, int16_t *pcm,
                                int64_t strideSamples, int64_t *nSamplesOut, int64_t *hdrNumSamples, int32_t *hdrSampleRate) {
    if (!ctx) return PAC_E_ARG;
    if (!pacOff || S <= 0) FAIL(PAC_E_ARG, "bad arguments to pac_decode_batch");
    std::vector<int64_t> len(S);
    for (int s = 0; s < S; s++) len[s] = pacOff[s + 1] - pacOff[s];
    return pac_decode_batch_strided(ctx, pac, pacOff, len.data(), S, pcm, strideSamples, nSamplesOut, hdrNumSamples, hdrSampleRate);
}

// ------------------------------------------------------------------ window + MDCT stage by itself (whole streams)
template <typename T>
static int mdct_batch_t(PacCtx *ctx, const int16_t *pcm, int64_t stride, const int64_t *nSamples, int S, double *lines, int32_t *oscale,
                        double *deviceMs, bool full = false) {
    const int M = ctx->M;
    const bool pcmDev = is_device_ptr(pcm);
    int64_t maxBlocks = 0;
    for (int s = 0; s < S; s++) {
        if (nSamples[s] < 0 || nSamples[s] > stride) FAIL(PAC_E_ARG, "nSamples[%d] outside [0, strideSamples]", s);
        const int64_t nb = pac_num_blocks(ctx, nSamples[s]);
        if (nb > maxBlocks) maxBlocks = nb;
    }
    const int16_t *d_pcm = pcm;
    if (!pcmDev) {
        CK(ctx->w_pcm.ensure((size_t)S * stride * 4 + 16));
        CK(cudaMemcpyAsync(ctx->w_pcm.p, pcm, (size_t)S * stride * 4, cudaMemcpyHostToDevice, ctx->stream));
        d_pcm = ctx->w_pcm.as<int16_t>();
    }
    CK(ctx->w_ns.ensure((size_t)S * 8));
    CK(cudaMemcpyAsync(ctx->w_ns.p, nSamples, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int TB = (int)(((int64_t)1 << 18) / S);                   // the encoder's tile budget: 2^18 (stream, block) items
    if (TB < 1) TB = 1;
    if (TB > maxBlocks) TB = (int)maxBlocks;
    const int64_t nworkMax = (int64_t)S * TB;
    CK(ctx->w_lines.ensure((size_t)nworkMax * 2 * M * sizeof(T)));
    CK(ctx->w_osc.ensure((size_t)nworkMax * 2));
    if (full) {
        CK(ctx->w_smr.ensure((size_t)nworkMax * 2 * kMaxBands * sizeof(T)));
        CK(ctx->w_bmax.ensure((size_t)nworkMax * 2 * kMaxBands * sizeof(T)));
        CK(ctx->w_lrms.ensure((size_t)nworkMax * 4));
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float total = 0.f;
    std::vector<T> hl;
    std::vector<uint8_t> ho;
    for (int64_t b0 = 0; b0 < maxBlocks; b0 += TB) {
        const int nb = (int)(maxBlocks - b0 < TB ? maxBlocks - b0 : TB);
        AnalysisArgs<T> aa{};
        aa.pcm = d_pcm; aa.strideSamples = stride; aa.nSamples = ctx->w_ns.as<int64_t>(); aa.blocks = nullptr;
        aa.S = S; aa.b0 = (int)b0; aa.nb = nb; aa.nwork = (int64_t)S * nb;
        aa.lines = ctx->w_lines.as<T>(); aa.oscale = ctx->w_osc.as<uint8_t>();
        aa.bands = ctx->bands; aa.nScaleBits = ctx->p.nScaleBits;
        int rc = get_tables<T>(ctx, ctx->N, &aa.tab);
        if (rc) return rc;
        if ((rc = get_tables<double>(ctx, ctx->N, &aa.tabd))) return rc;
        if (full) { aa.smr = ctx->w_smr.as<T>(); aa.bmax = ctx->w_bmax.as<T>(); aa.lrms = ctx->w_lrms.as<uint32_t>(); }
        CK(cudaEventRecord(e0, ctx->stream));
        if (full) rc = launch_analysis<T>(ctx, aa);           // the whole analysis stage, by itself (persistent grid)
        else if constexpr (sizeof(T) == 4) rc = launch_mdct(ctx, aa);               // fp32 mode: the stand-alone k_mdct
        else if (ctx->LOGM == 10) rc = launch_mdct_only_t<double, 10>(ctx, aa);     // fp64 mode: sections A + C of k_analysis
        else rc = launch_mdct_only_t<double, 9>(ctx, aa);
        if (rc) return rc;
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        total += ms;
        if (lines || oscale) {
            hl.resize((size_t)aa.nwork * 2 * M); ho.resize((size_t)aa.nwork * 2);
            CK(cudaMemcpy(hl.data(), aa.lines, hl.size() * sizeof(T), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(ho.data(), aa.oscale, ho.size(), cudaMemcpyDeviceToHost));
            for (int s = 0; s < S; s++) {
                const int64_t nbs = pac_num_blocks(ctx, nSamples[s]);
                for (int64_t b = b0; b < b0 + nb && b < nbs; b++) {
                    const int64_t w = (int64_t)s * nb + (b - b0), gi = (int64_t)s * maxBlocks + b;
                    for (int ch = 0; ch < 2; ch++) {
                        if (oscale) oscale[gi * 2 + ch] = ho[w * 2 + ch];
                        if (lines) for (int i = 0; i < M; i++) lines[(gi * 2 + ch) * M + i] = (double)hl[(w * 2 + ch) * M + i];
                    }
                }
            }
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (deviceMs) *deviceMs = (double)total;
    return PAC_OK;
}

extern "C" int pac_mdct_batch(PacCtx *ctx, const int16_t *pcm, int64_t strideSamples, const int64_t *nSamples, int S, double *lines,
                              int32_t *oscale, double *deviceMs) {
    if (!ctx) return PAC_E_ARG;
    if (!pcm || !nSamples || S <= 0 || strideSamples < 0) FAIL(PAC_E_ARG, "bad arguments to pac_mdct_batch");
    CK(cudaSetDevice(ctx->device));
    if (ctx->precision == PAC_PRECISION_FP64) return mdct_batch_t<double>(ctx, pcm, strideSamples, nSamples, S, lines, oscale, deviceMs);
    return mdct_batch_t<float>(ctx, pcm, strideSamples, nSamples, S, lines, oscale, deviceMs);
}

extern "C" int pac_analysis_batch(PacCtx *ctx, const int16_t *pcm, int64_t strideSamples, const int64_t *nSamples, int S, double *deviceMs) {
    if (!ctx) return PAC_E_ARG;
    if (!pcm || !nSamples || S <= 0 || strideSamples < 0) FAIL(PAC_E_ARG, "bad arguments to pac_analysis_batch");
    CK(cudaSetDevice(ctx->device));
    if (ctx->precision == PAC_PRECISION_FP64) return mdct_batch_t<double>(ctx, pcm, strideSamples, nSamples, S, nullptr, nullptr, deviceMs, true);
    return mdct_batch_t<float>(ctx, pcm, strideSamples, nSamples, S, nullptr, nullptr, deviceMs, true);
}

// ------------------------------------------------------------------ per-block API
template <typename T>
static int encode_blocks_t(PacCtx *ctx, const double *data, int nblk, PacStreamState *state, int32_t *scaleFactor, int32_t *bitAlloc,
                           int32_t *mant, int32_t *tableID, int32_t *overallScale, int32_t *lrms, uint8_t *chunk, int64_t chunkCap,
                           int32_t *chunkBytes) {
    const int M = ctx->M, N = ctx->N, NB = ctx->bands.nBands;
    const int64_t nwork = nblk;
    CK(ctx->w_misc.ensure((size_t)nwork * 2 * N * 8));
    CK(cudaMemcpyAsync(ctx->w_misc.p, data, (size_t)nwork * 2 * N * 8, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<StreamState> st(nblk);
    for (int i = 0; i < nblk; i++) { st[i].extraBits = state[i].extraBits; st[i].bitDeposit = state[i].bitDeposit; st[i].outOffset = 0; st[i].reserved = 0; }
    CK(ctx->w_state.ensure((size_t)nblk * sizeof(StreamState)));
    CK(cudaMemcpyAsync(ctx->w_state.p, st.data(), (size_t)nblk * sizeof(StreamState), cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->w_lines.ensure((size_t)nwork * 2 * M * sizeof(T)));
    CK(ctx->w_smr.ensure((size_t)nwork * 2 * kMaxBands * sizeof(T)));
    CK(ctx->w_bmax.ensure((size_t)nwork * 2 * kMaxBands * sizeof(T)));
    CK(ctx->w_osc.ensure((size_t)nwork * 2));
    CK(ctx->w_lrms.ensure((size_t)nwork * 4));
    CK(ctx->w_ba.ensure((size_t)nwork * 2 * kMaxBands));
    CK(ctx->w_sf.ensure((size_t)nwork * 2 * kMaxBands));
    CK(ctx->w_tid.ensure((size_t)nwork * 2));
    CK(ctx->w_nby.ensure((size_t)nwork * 2 * 4));
    CK(ctx->w_coff.ensure((size_t)nwork * 2 * 8));
    const int64_t ccap = 4096;
    CK(ctx->w_out.ensure((size_t)nwork * 2 * ccap));
    CK(ctx->w_misc2.ensure((size_t)nwork * 2 * M * 4));
    CK(cudaMemsetAsync(ctx->w_misc2.p, 0, (size_t)nwork * 2 * M * 4, ctx->stream));
    CK(ctx->w_ovf.ensure((size_t)nblk * 4));
    CK(cudaMemsetAsync(ctx->w_ovf.p, 0, (size_t)nblk * 4, ctx->stream));
    AnalysisArgs<T> aa{};
    aa.pcm = nullptr; aa.blocks = ctx->w_misc.as<double>(); aa.S = nblk; aa.b0 = 0; aa.nb = 1; aa.nwork = nwork;
    aa.lines = ctx->w_lines.as<T>(); aa.smr = ctx->w_smr.as<T>(); aa.bmax = ctx->w_bmax.as<T>();
    aa.oscale = ctx->w_osc.as<uint8_t>(); aa.lrms = ctx->w_lrms.as<uint32_t>();
    int rc = launch_analysis<T>(ctx, aa);
    if (rc) return rc;
    ScanArgs<T> sa{};
    sa.S = nblk; sa.b0 = 0; sa.nb = 1; sa.nSamples = nullptr; sa.state = ctx->w_state.as<StreamState>();
    sa.lines = aa.lines; sa.smr = aa.smr; sa.bmax = aa.bmax; sa.lrms = aa.lrms;
    sa.ba = ctx->w_ba.as<uint8_t>(); sa.sf = ctx->w_sf.as<uint8_t>(); sa.tableID = ctx->w_tid.as<uint8_t>();
    sa.nbytes = ctx->w_nby.as<uint32_t>(); sa.chunkOff = ctx->w_coff.as<long long>();
    if ((rc = launch_scan<T>(ctx, sa))) return rc;
    PackArgs<T> pa{};
    pa.S = nblk; pa.b0 = 0; pa.nb = 1; pa.nSamples = nullptr;
    pa.lines = aa.lines; pa.ba = sa.ba; pa.sf = sa.sf; pa.tableID = sa.tableID; pa.oscale = aa.oscale; pa.lrms = aa.lrms;
    pa.nbytes = sa.nbytes; pa.chunkOff = sa.chunkOff; pa.out = ctx->w_out.as<uint8_t>(); pa.cap = ccap; pa.perChunk = 1;
    pa.overflow = ctx->w_ovf.as<int>(); pa.o_mant = ctx->w_misc2.as<int32_t>(); pa.header = nullptr; pa.headerBytes = 0;
    if ((rc = launch_pack<T>(ctx, pa))) return rc;
    std::vector<uint8_t> hba((size_t)nwork * 2 * kMaxBands), hsf((size_t)nwork * 2 * kMaxBands), htid((size_t)nwork * 2), hosc((size_t)nwork * 2);
    std::vector<uint32_t> hlr(nwork), hnby((size_t)nwork * 2);
    CK(cudaMemcpyAsync(hba.data(), sa.ba, hba.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hsf.data(), sa.sf, hsf.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(htid.data(), sa.tableID, htid.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hosc.data(), aa.oscale, hosc.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hlr.data(), aa.lrms, hlr.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hnby.data(), sa.nbytes, hnby.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(st.data(), ctx->w_state.p, (size_t)nblk * sizeof(StreamState), cudaMemcpyDeviceToHost, ctx->stream));
    if (mant) CK(cudaMemcpyAsync(mant, ctx->w_misc2.p, (size_t)nwork * 2 * M * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int64_t w = 0; w < nwork; w++) {
        state[w].extraBits = st[w].extraBits; state[w].bitDeposit = st[w].bitDeposit;
        if (lrms) lrms[w] = (int32_t)hlr[w];
        for (int ch = 0; ch < 2; ch++) {
            if (tableID) tableID[w * 2 + ch] = htid[w * 2 + ch];
            if (overallScale) overallScale[w * 2 + ch] = hosc[w * 2 + ch];
            if (chunkBytes) chunkBytes[w * 2 + ch] = (int32_t)hnby[w * 2 + ch];
            for (int bd = 0; bd < NB; bd++) {
                if (scaleFactor) scaleFactor[(w * 2 + ch) * NB + bd] = hsf[(w * 2 + ch) * kMaxBands + bd];
                if (bitAlloc) bitAlloc[(w * 2 + ch) * NB + bd] = hba[(w * 2 + ch) * kMaxBands + bd];
            }
            if (chunk) {
                if ((int64_t)hnby[w * 2 + ch] > chunkCap) FAIL(PAC_E_OVERFLOW, "chunkCap too small (%u needed)", hnby[w * 2 + ch]);
                CK(cudaMemcpy(chunk + (w * 2 + ch) * chunkCap, ctx->w_out.as<uint8_t>() + (w * 2 + ch) * ccap, hnby[w * 2 + ch], cudaMemcpyDeviceToHost));
            }
        }
    }
    return PAC_OK;
}

extern "C" int pac_encode_blocks(PacCtx *ctx, const double *data, int nblk, PacStreamState *state, int32_t *scaleFactor,
                                 int32_t *bitAlloc, int32_t *mant, int32_t *tableID, int32_t *overallScale, int32_t *lrms,
                                 uint8_t *chunk, int64_t chunkCap, int32_t *chunkBytes) {
    if (!ctx) return PAC_E_ARG;
    if (!data || !state || nblk <= 0) FAIL(PAC_E_ARG, "bad arguments to pac_encode_blocks");
    CK(cudaSetDevice(ctx->device));
    if (ctx->precision == PAC_PRECISION_FP64)
        return encode_blocks_t<double>(ctx, data, nblk, state, scaleFactor, bitAlloc, mant, tableID, overallScale, lrms, chunk, chunkCap, chunkBytes);
    return encode_blocks_t<float>(ctx, data, nblk, state, scaleFactor, bitAlloc, mant, tableID, overallScale, lrms, chunk, chunkCap, chunkBytes);
}

// dequantise kernel for the per-block decode API (codec.py:31-43)
template <typename T>
__global__ void k_dequant_blocks(const int32_t *sf, const int32_t *ba, const int32_t *mant, const int32_t *oscale, int nblk, int M,
                                 int nScaleBits, T *lines, BandInfo bands) {
    const int NB = bands.nBands;
    int64_t total = (int64_t)nblk * 2 * M;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t c = e / M;
        int i = (int)(e - c * M);
        int bd = 0;
        while (bd < NB - 1 && i >= bands.lo[bd + 1]) bd++;
        int b = ba[c * NB + bd];
        double v = 0.0;
        if (b) v = dequant(sf[c * NB + bd], (long long)mant[e], (1 << nScaleBits) - 1, b) / (double)(1 << oscale[c]);
        lines[e] = (T)v;
    }
}

template <typename T>
static int decode_blocks_t(PacCtx *ctx, const int32_t *scaleFactor, const int32_t *bitAlloc, const int32_t *mant,
                           const int32_t *overallScale, const int32_t *lrms, int nblk, double *out) {
    const int M = ctx->M, N = ctx->N, NB = ctx->bands.nBands;
    CK(ctx->w_misc.ensure((size_t)nblk * 2 * NB * 4));
    CK(ctx->w_misc2.ensure((size_t)nblk * 2 * NB * 4));
    CK(ctx->w_misc3.ensure((size_t)nblk * 2 * M * 4));
    CK(ctx->w_misc4.ensure((size_t)nblk * 2 * 4));
    CK(ctx->w_lrms.ensure((size_t)nblk * 4));
    CK(ctx->w_lines.ensure((size_t)nblk * 2 * M * sizeof(T)));
    CK(ctx->w_misc5.ensure((size_t)nblk * 2 * N * 8));
    CK(cudaMemcpyAsync(ctx->w_misc.p, scaleFactor, (size_t)nblk * 2 * NB * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w_misc2.p, bitAlloc, (size_t)nblk * 2 * NB * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w_misc3.p, mant, (size_t)nblk * 2 * M * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w_misc4.p, overallScale, (size_t)nblk * 2 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w_lrms.p, lrms, (size_t)nblk * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_dequant_blocks<T><<<(unsigned)((nblk * 2 * M + 255) / 256), 256, 0, ctx->stream>>>(
        ctx->w_misc.as<int32_t>(), ctx->w_misc2.as<int32_t>(), ctx->w_misc3.as<int32_t>(), ctx->w_misc4.as<int32_t>(), nblk, M,
        ctx->p.nScaleBits, ctx->w_lines.as<T>(), ctx->bands);
    ctx->launches++;
    CK(cudaGetLastError());
    SynthArgs<T> sa{};
    sa.lines = ctx->w_lines.as<T>(); sa.lrms = ctx->w_lrms.as<uint32_t>(); sa.nBlocks = nullptr; sa.S = 1; sa.maxBlocks = nblk; sa.run = 1;
    sa.pcm = nullptr; sa.strideSamples = 0; sa.nSamplesOut = nullptr; sa.rawOut = ctx->w_misc5.as<double>();
    int rc = launch_synth<T>(ctx, sa, nblk);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, ctx->w_misc5.p, (size_t)nblk * 2 * N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PAC_OK;
}

extern "C" int pac_decode_blocks(PacCtx *ctx, const int32_t *scaleFactor, const int32_t *bitAlloc, const int32_t *mant,
                                 const int32_t *overallScale, const int32_t *lrms, int nblk, double *out) {
    if (!ctx) return PAC_E_ARG;
    if (!scaleFactor || !bitAlloc || !mant || !overallScale || !lrms || !out || nblk <= 0) FAIL(PAC_E_ARG, "bad arguments to pac_decode_blocks");
    CK(cudaSetDevice(ctx->device));
    if (ctx->precision == PAC_PRECISION_FP64) return decode_blocks_t<double>(ctx, scaleFactor, bitAlloc, mant, overallScale, lrms, nblk, out);
    return decode_blocks_t<float>(ctx, scaleFactor, bitAlloc, mant, overallScale, lrms, nblk, out);
}

extern "C" int pac_unpack_blocks(PacCtx *ctx, const uint8_t *chunks, int64_t chunkCap, const int32_t *chunkBytes, int nblk,
                                 int32_t *scaleFactor, int32_t *bitAlloc, int32_t *mant, int32_t *overallScale, int32_t *lrms,
                                 int32_t *tableID) {
    if (!ctx) return PAC_E_ARG;
    if (!chunks || !chunkBytes || nblk <= 0) FAIL(PAC_E_ARG, "bad arguments to pac_unpack_blocks");
    CK(cudaSetDevice(ctx->device));
    const int M = ctx->M, NB = ctx->bands.nBands;
    const int64_t nchunk = (int64_t)nblk * 2;
    CK(ctx->w_misc.ensure((size_t)nchunk * chunkCap + 16));
    CK(cudaMemcpyAsync(ctx->w_misc.p, chunks, (size_t)nchunk * chunkCap, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int64_t> pos(nchunk);
    for (int64_t c = 0; c < nchunk; c++) pos[c] = c * chunkCap;
    CK(ctx->w_coff.ensure((size_t)nchunk * 8));
    CK(cudaMemcpyAsync(ctx->w_coff.p, pos.data(), (size_t)nchunk * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->w_nby.ensure((size_t)nchunk * 4));
    CK(cudaMemcpyAsync(ctx->w_nby.p, chunkBytes, (size_t)nchunk * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->w_lrms.ensure((size_t)nblk * 4));
    CK(ctx->w_misc2.ensure((size_t)nchunk * kMaxBands * 4));
    CK(ctx->w_misc3.ensure((size_t)nchunk * kMaxBands * 4));
    CK(ctx->w_misc4.ensure((size_t)nchunk * M * 4));
    CK(ctx->w_misc5.ensure((size_t)nchunk * 4));
    CK(ctx->w_misc6.ensure((size_t)nchunk * 4));
    CK(ctx->w_ovf.ensure(4));
    CK(cudaMemsetAsync(ctx->w_ovf.p, 0, 4, ctx->stream));
    UnpackArgs<double> ua{};
    ua.pac = ctx->w_misc.as<uint8_t>(); ua.chunkPos = ctx->w_coff.as<int64_t>(); ua.chunkLen = ctx->w_nby.as<int32_t>(); ua.nBlocks = nullptr;
    ua.S = 1; ua.maxBlocks = nblk; ua.M = M;
    ua.nScaleBits = ctx->p.nScaleBits; ua.nMantSizeBits = ctx->p.nMantSizeBits; ua.nTableIDBits = 4;
    ua.codes = nullptr; ua.meta = nullptr; ua.oscale = nullptr; ua.lrms = ctx->w_lrms.as<uint32_t>(); ua.err = ctx->w_ovf.as<int32_t>();
    ua.o_sf = ctx->w_misc2.as<int32_t>(); ua.o_ba = ctx->w_misc3.as<int32_t>(); ua.o_mant = ctx->w_misc4.as<int32_t>();
    ua.o_oscale = ctx->w_misc5.as<int32_t>(); ua.o_tableID = ctx->w_misc6.as<int32_t>();
    ua.dt = ctx->dt; ua.bands = ctx->bands;
    k_unpack<double><<<(unsigned)((nchunk + kUnpackThreads - 1) / kUnpackThreads), kUnpackThreads, 0, ctx->stream>>>(ua);
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<int32_t> hsf((size_t)nchunk * kMaxBands), hba((size_t)nchunk * kMaxBands);
    int32_t err = 0;
    CK(cudaMemcpyAsync(hsf.data(), ua.o_sf, hsf.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hba.data(), ua.o_ba, hba.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (mant) CK(cudaMemcpyAsync(mant, ua.o_mant, (size_t)nchunk * M * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (overallScale) CK(cudaMemcpyAsync(overallScale, ua.o_oscale, (size_t)nchunk * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (tableID) CK(cudaMemcpyAsync(tableID, ua.o_tableID, (size_t)nchunk * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (lrms) CK(cudaMemcpyAsync(lrms, ua.lrms, (size_t)nblk * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&err, ua.err, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int64_t c = 0; c < nchunk; c++)
        for (int bd = 0; bd < NB; bd++) {
            if (scaleFactor) scaleFactor[c * NB + bd] = hsf[c * kMaxBands + bd];
            if (bitAlloc) bitAlloc[c * NB + bd] = hba[c * kMaxBands + bd];
        }
    if (err) FAIL(PAC_E_FORMAT, "malformed chunk (bad table ID or code)");
    return PAC_OK;
}

// ------------------------------------------------------------------ L2 entry points
static int get_window(PacCtx *ctx, int kind, int N, const double **dptr) {
    int key = kind * 65536 + ilog2(N);
    auto it = ctx->winTables.find(key);
    if (it == ctx->winTables.end()) {
        std::vector<double> w(N);
        double Nf = (double)N;
        if (kind == 0) for (int n = 0; n < N; n++) w[n] = sin((n + 0.5) * M_PI / Nf);                        // window.py:35-37
        else if (kind == 1) for (int n = 0; n < N; n++) w[n] = 0.5 * (1 - cos(2.0 * (n + 0.5) * M_PI / Nf)); // window.py:49-51
        else kbd_window_host(N, w.data());                                                                  // window.py:56-78, alpha = 4
        void *d = nullptr;
        CK(cudaMalloc(&d, (size_t)N * 8));
        CK(cudaMemcpy(d, w.data(), (size_t)N * 8, cudaMemcpyHostToDevice));
        it = ctx->winTables.emplace(key, d).first;
    }
    *dptr = reinterpret_cast<const double *>(it->second);
    return PAC_OK;
}

extern "C" int pac_window(PacCtx *ctx, int kind, double *x, int n, int N) {
    if (!ctx) return PAC_E_ARG;
    if (!x || n <= 0 || N <= 0 || kind < 0 || kind > 2 || (kind == 2 && (N & 1))) FAIL(PAC_E_ARG, "bad arguments to pac_window");
    CK(cudaSetDevice(ctx->device));
    const double *w;
    int rc = get_window(ctx, kind, N, &w);
    if (rc) return rc;
    int64_t total = (int64_t)n * N;
    CK(ctx->w_misc.ensure((size_t)total * 8));
    CK(cudaMemcpyAsync(ctx->w_misc.p, x, (size_t)total * 8, cudaMemcpyHostToDevice, ctx->stream));
    int64_t grid = (total + 255) / 256;
    if (grid > 65535) grid = 65535;
    k_window<<<(unsigned)grid, 256, 0, ctx->stream>>>(ctx->w_misc.as<double>(), w, total, N);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(x, ctx->w_misc.p, (size_t)total * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PAC_OK;
}

template <typename T, int LOGM>
static int mdct_run(PacCtx *ctx, bool inverse, const double *in, int n, double *out) {
    constexpr int M = 1 << LOGM, N = 2 * M, NT = (M / 4 > 0 ? M / 4 : 1);
    DevTables<T> tb;
    int rc = get_tables<T>(ctx, N, &tb);
    if (rc) return rc;
    size_t nin = (size_t)n * (inverse ? M : N) * 8, nout = (size_t)n * (inverse ? N : M) * 8;
    CK(ctx->w_misc.ensure(nin));
    CK(ctx->w_misc2.ensure(nout));
    CK(cudaMemcpyAsync(ctx->w_misc.p, in, nin, cudaMemcpyHostToDevice, ctx->stream));
    size_t smem = sizeof(MdctSmem<T, LOGM>);
    if (inverse) {
        CK(cudaFuncSetAttribute(k_imdct<T, LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_imdct<T, LOGM><<<n, NT, smem, ctx->stream>>>(ctx->w_misc.as<double>(), ctx->w_misc2.as<double>(), n, tb);
    } else {
        CK(cudaFuncSetAttribute(k_mdct<T, LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_mdct<T, LOGM><<<n, NT, smem, ctx->stream>>>(ctx->w_misc.as<double>(), ctx->w_misc2.as<double>(), n, tb);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, ctx->w_misc2.p, nout, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PAC_OK;
}

template <typename T>
static int mdct_dispatch(PacCtx *ctx, bool inverse, const double *in, int n, int N, double *out) {
    switch (N) {
        case 8: return mdct_run<T, 2>(ctx, inverse, in, n, out);
        case 16: return mdct_run<T, 3>(ctx, inverse, in, n, out);
        case 256: return mdct_run<T, 7>(ctx, inverse, in, n, out);
        case 512: return mdct_run<T, 8>(ctx, inverse, in, n, out);
        case 1024: return mdct_run<T, 9>(ctx, inverse, in, n, out);
        case 2048: return mdct_run<T, 10>(ctx, inverse, in, n, out);
        case 4096: return mdct_run<T, 11>(ctx, inverse, in, n, out);
    }
    FAIL(PAC_E_ARG, "MDCT block length %d not supported (8, 16, 256, 512, 1024, 2048, 4096)", N);
}

extern "C" int pac_mdct(PacCtx *ctx, const double *x, int n, int N, double *X) {
    if (!ctx) return PAC_E_ARG;
    if (!x || !X || n <= 0) FAIL(PAC_E_ARG, "bad arguments to pac_mdct");
    CK(cudaSetDevice(ctx->device));
    return ctx->precision == PAC_PRECISION_FP64 ? mdct_dispatch<double>(ctx, false, x, n, N, X) : mdct_dispatch<float>(ctx, false, x, n, N, X);
}

extern "C" int pac_imdct(PacCtx *ctx, const double *X, int n, int N, double *x) {
    if (!ctx) return PAC_E_ARG;
    if (!x || !X || n <= 0) FAIL(PAC_E_ARG, "bad arguments to pac_imdct");
    CK(cudaSetDevice(ctx->device));
    return ctx->precision == PAC_PRECISION_FP64 ? mdct_dispatch<double>(ctx, true, X, n, N, x) : mdct_dispatch<float>(ctx, true, X, n, N, x);
}

template <typename T>
static int analysis_t(PacCtx *ctx, const double *data, int nblk, int32_t *lrms, int32_t *oscale, double *mdct, double *bthr,
                      double *smr, double *lines) {
    const int M = ctx->M, N = ctx->N, NB = ctx->bands.nBands;
    const int64_t nwork = nblk;
    CK(ctx->w_misc.ensure((size_t)nwork * 2 * N * 8));
    CK(cudaMemcpyAsync(ctx->w_misc.p, data, (size_t)nwork * 2 * N * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx->w_lines.ensure((size_t)nwork * 2 * M * sizeof(T)));
    CK(ctx->w_smr.ensure((size_t)nwork * 2 * kMaxBands * sizeof(T)));
    CK(ctx->w_bmax.ensure((size_t)nwork * 2 * kMaxBands * sizeof(T)));
    CK(ctx->w_osc.ensure((size_t)nwork * 2));
    CK(ctx->w_lrms.ensure((size_t)nwork * 4));
    CK(ctx->w_dbg1.ensure((size_t)nwork * 2 * M * sizeof(T)));
    CK(ctx->w_dbg2.ensure((size_t)nwork * 6 * M * sizeof(T)));
    AnalysisArgs<T> aa{};
    aa.pcm = nullptr; aa.blocks = ctx->w_misc.as<double>(); aa.S = nblk; aa.b0 = 0; aa.nb = 1; aa.nwork = nwork;
    aa.lines = ctx->w_lines.as<T>(); aa.smr = ctx->w_smr.as<T>(); aa.bmax = ctx->w_bmax.as<T>();
    aa.oscale = ctx->w_osc.as<uint8_t>(); aa.lrms = ctx->w_lrms.as<uint32_t>();
    aa.dbg_mdct = ctx->w_dbg1.as<T>(); aa.dbg_bthr = ctx->w_dbg2.as<T>();
    int rc = launch_analysis<T>(ctx, aa);
    if (rc) return rc;
    std::vector<T> hl((size_t)nwork * 2 * M), hm((size_t)nwork * 2 * M), hb((size_t)nwork * 6 * M), hs((size_t)nwork * 2 * kMaxBands);
    std::vector<uint8_t> hosc((size_t)nwork * 2);
    std::vector<uint32_t> hlr(nwork);
    CK(cudaMemcpyAsync(hl.data(), aa.lines, hl.size() * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hm.data(), aa.dbg_mdct, hm.size() * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hb.data(), aa.dbg_bthr, hb.size() * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hs.data(), aa.smr, hs.size() * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hosc.data(), aa.oscale, hosc.size(), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(hlr.data(), aa.lrms, hlr.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int64_t w = 0; w < nwork; w++) {
        if (lrms) lrms[w] = (int32_t)hlr[w];
        for (int ch = 0; ch < 2; ch++) {
            if (oscale) oscale[w * 2 + ch] = hosc[w * 2 + ch];
            if (smr) for (int bd = 0; bd < NB; bd++) smr[(w * 2 + ch) * NB + bd] = (double)hs[(w * 2 + ch) * kMaxBands + bd];
        }
    }
    if (lines) for (size_t i = 0; i < hl.size(); i++) lines[i] = (double)hl[i];
    if (mdct) for (size_t i = 0; i < hm.size(); i++) mdct[i] = (double)hm[i];
    if (bthr) for (size_t i = 0; i < hb.size(); i++) bthr[i] = (double)hb[i];
    return PAC_OK;
}

extern "C" int pac_analysis(PacCtx *ctx, const double *data, int nblk, int32_t *lrms, int32_t *oscale, double *mdct, double *bthr,
                            double *smr, double *lines) {
    if (!ctx) return PAC_E_ARG;
    if (!data || nblk <= 0) FAIL(PAC_E_ARG, "bad arguments to pac_analysis");
    CK(cudaSetDevice(ctx->device));
    if (ctx->precision == PAC_PRECISION_FP64) return analysis_t<double>(ctx, data, nblk, lrms, oscale, mdct, bthr, smr, lines);
    return analysis_t<float>(ctx, data, nblk, lrms, oscale, mdct, bthr, smr, lines);
}

template <typename T, int LOGM>
static int calc_smrs_t(PacCtx *ctx, const double *data, const double *mdct, int n, int scale, double *smr, int noDrop, double *thr) {
    constexpr int M = 1 << LOGM, N = 2 * M;
    const int NB = ctx->bands.nBands;
    CK(ctx->w_misc.ensure((size_t)n * N * 8));
    CK(ctx->w_misc2.ensure((size_t)n * M * 8));
    CK(ctx->w_misc3.ensure((size_t)n * kMaxBands * 8));
    CK(ctx->w_misc4.ensure((size_t)n * M * 8));
    CK(cudaMemcpyAsync(ctx->w_misc.p, data, (size_t)n * N * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (mdct) CK(cudaMemcpyAsync(ctx->w_misc2.p, mdct, (size_t)n * M * 8, cudaMemcpyHostToDevice, ctx->stream));
    SmrMonoArgs<T> a{};
    a.data = ctx->w_misc.as<double>(); a.mdct = ctx->w_misc2.as<double>(); a.n = n; a.scale = scale; a.noDrop = noDrop;
    a.smr = smr ? ctx->w_misc3.as<double>() : nullptr; a.thr = thr ? ctx->w_misc4.as<double>() : nullptr; a.bands = ctx->bands;
    int rc = get_tables<T>(ctx, N, &a.tab);
    if (rc) return rc;
    if ((rc = get_tables<double>(ctx, N, &a.tabd))) return rc;
    size_t smem = sizeof(AnalysisSmem<T, LOGM>);
    CK(cudaFuncSetAttribute(k_calc_smrs<T, LOGM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_calc_smrs<T, LOGM><<<n, M / 4, smem, ctx->stream>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<double> hs((size_t)n * kMaxBands);
    if (smr) CK(cudaMemcpyAsync(hs.data(), a.smr, hs.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (thr) CK(cudaMemcpyAsync(thr, a.thr, (size_t)n * M * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (smr) for (int i = 0; i < n; i++) for (int bd = 0; bd < NB; bd++) smr[(size_t)i * NB + bd] = hs[(size_t)i * kMaxBands + bd];
    return PAC_OK;
}

static int calc_smrs_dispatch(PacCtx *ctx, const double *data, const double *mdct, int n, int scale, double *smr, int noDrop, double *thr) {
    const bool f64 = ctx->precision == PAC_PRECISION_FP64;
    if (ctx->LOGM == 10) return f64 ? calc_smrs_t<double, 10>(ctx, data, mdct, n, scale, smr, noDrop, thr) : calc_smrs_t<float, 10>(ctx, data, mdct, n, scale, smr, noDrop, thr);
    if (ctx->LOGM == 9) return f64 ? calc_smrs_t<double, 9>(ctx, data, mdct, n, scale, smr, noDrop, thr) : calc_smrs_t<float, 9>(ctx, data, mdct, n, scale, smr, noDrop, thr);
    FAIL(PAC_E_ARG, "unsupported nMDCTLines");
}

extern "C" int pac_masked_threshold(PacCtx *ctx, const double *data, int n, int noDrop, double *thr) {
    if (!ctx) return PAC_E_ARG;
    if (!data || !thr || n <= 0) FAIL(PAC_E_ARG, "bad arguments to pac_masked_threshold");
    CK(cudaSetDevice(ctx->device));
    return calc_smrs_dispatch(ctx, data, nullptr, n, 0, nullptr, noDrop, thr);
}

extern "C" int pac_huffman_select(PacCtx *ctx, const uint32_t *mag, const int32_t *ba, int n, int32_t *tableID, int64_t *totals) {
    if (!ctx) return PAC_E_ARG;
    if (n < 0 || !tableID || !totals || (n > 0 && (!mag || !ba))) FAIL(PAC_E_ARG, "bad arguments to pac_huffman_select");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w_misc.ensure((size_t)(n + 1) * 4)); CK(ctx->w_misc2.ensure((size_t)(n + 1) * 4));
    CK(ctx->w_misc3.ensure(4)); CK(ctx->w_misc4.ensure(kNTables * 8));
    if (n) {
        CK(cudaMemcpyAsync(ctx->w_misc.p, mag, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->w_misc2.p, ba, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    }
    k_huff_select<<<1, 256, 0, ctx->stream>>>(ctx->w_misc.as<uint32_t>(), ctx->w_misc2.as<int32_t>(), n, ctx->lenLut, ctx->ec,
                                            ctx->w_misc3.as<int32_t>(), ctx->w_misc4.as<long long>());
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(tableID, ctx->w_misc3.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(totals, ctx->w_misc4.p, kNTables * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PAC_OK;
}

extern "C" int pac_calc_smrs(PacCtx *ctx, const double *data, const double *mdct, int n, int scale, double *smr) {
    if (!ctx) return PAC_E_ARG;
    if (!data || !mdct || !smr || n <= 0) FAIL(PAC_E_ARG, "bad arguments to pac_calc_smrs");
    CK(cudaSetDevice(ctx->device));
    return calc_smrs_dispatch(ctx, data, mdct, n, scale, smr, 0, nullptr);
}

extern "C" int pac_bitalloc(PacCtx *ctx, int n, const double *bitBudget, const int64_t *extraBits, int maxMantBits,
                            const double *smr, const int32_t *lrms, int32_t *bits, int64_t *bitDifference) {
    if (!ctx) return PAC_E_ARG;
    if (n <= 0 || !bitBudget || !extraBits || !smr || !lrms || !bits || !bitDifference) FAIL(PAC_E_ARG, "bad arguments to pac_bitalloc");
    CK(cudaSetDevice(ctx->device));
    const int NB = ctx->bands.nBands;
    CK(ctx->w_misc.ensure((size_t)n * 8)); CK(ctx->w_misc2.ensure((size_t)n * 8)); CK(ctx->w_misc3.ensure((size_t)n * NB * 8));
    CK(ctx->w_misc4.ensure((size_t)n * 4)); CK(ctx->w_misc5.ensure((size_t)n * NB * 4)); CK(ctx->w_misc6.ensure((size_t)n * 8));
    CK(cudaMemcpyAsync(ctx->w_misc.p, bitBudget, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w_misc2.p, extraBits, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w_misc3.p, smr, (size_t)n * NB * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->w_misc4.p, lrms, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    k_bitalloc<<<(n + 3) / 4, 128, 0, ctx->stream>>>(n, ctx->w_misc.as<double>(), ctx->w_misc2.as<long long>(), maxMantBits,
                                                     ctx->w_misc3.as<double>(), ctx->w_misc4.as<uint32_t>(), ctx->w_misc5.as<int32_t>(),
                                                     ctx->w_misc6.as<long long>(), ctx->bands);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(bits, ctx->w_misc5.p, (size_t)n * NB * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(bitDifference, ctx->w_misc6.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PAC_OK;
}

extern "C" int pac_bitalloc_alt(PacCtx *ctx, int mode, int n, const double *bitBudget, int maxMantBits, const double *level,
                                int32_t *bits) {
    if (!ctx) return PAC_E_ARG;
    if (n <= 0 || !bitBudget || !bits || mode < 0 || mode > 2 || (mode > 0 && !level)) FAIL(PAC_E_ARG, "bad arguments to pac_bitalloc_alt");
    CK(cudaSetDevice(ctx->device));
    const int NB = ctx->bands.nBands;
    CK(ctx->w_misc.ensure((size_t)n * 8)); CK(ctx->w_misc3.ensure((size_t)n * NB * 8));
    CK(ctx->w_misc4.ensure((size_t)n * 4)); CK(ctx->w_misc5.ensure((size_t)n * NB * 4));
    CK(cudaMemcpyAsync(ctx->w_misc.p, bitBudget, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    if (mode > 0) CK(cudaMemcpyAsync(ctx->w_misc3.p, level, (size_t)n * NB * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_bitalloc_alt<<<(n + 3) / 4, 128, 0, ctx->stream>>>(n, mode, ctx->w_misc.as<double>(), maxMantBits, ctx->w_misc3.as<double>(),
                                                         ctx->w_misc5.as<int32_t>(), ctx->w_misc4.as<int32_t>(), ctx->bands);
    ctx->launches++;
    CK(cudaGetLastError());
    std::vector<int32_t> st((size_t)n);
    CK(cudaMemcpyAsync(bits, ctx->w_misc5.p, (size_t)n * NB * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(st.data(), ctx->w_misc4.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < n; i++)
        if (st[i]) FAIL(PAC_E_ARG, "pac_bitalloc_alt: bits remain but no band can take one -- the reference loop (bitalloc.py:74,109) never terminates for this input");
    return PAC_OK;
}

extern "C" int pac_histogram(PacCtx *ctx, const uint32_t *codes, int64_t n, int64_t base, int nbins, int64_t *counts, int64_t *first) {
    if (!ctx) return PAC_E_ARG;
    if (n < 0 || nbins <= 0 || !counts || !first || (n > 0 && !codes)) FAIL(PAC_E_ARG, "bad arguments to pac_histogram");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->w_misc2.ensure((size_t)nbins * 8)); CK(ctx->w_misc3.ensure((size_t)nbins * 8));
    CK(cudaMemsetAsync(ctx->w_misc2.p, 0, (size_t)nbins * 8, ctx->stream));
    CK(cudaMemsetAsync(ctx->w_misc3.p, 0xff, (size_t)nbins * 8, ctx->stream));
    if (n > 0) {
        const uint32_t *d_codes = codes;
        cudaPointerAttributes at;
        bool onDevice = cudaPointerGetAttributes(&at, codes) == cudaSuccess && at.type == cudaMemoryTypeDevice;
        cudaGetLastError();
        if (!onDevice) {
            CK(ctx->w_misc.ensure((size_t)n * 4));
            CK(cudaMemcpyAsync(ctx->w_misc.p, codes, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
            d_codes = ctx->w_misc.as<uint32_t>();
        }
        int64_t grid = (n + 256 * 64 - 1) / (256 * 64);
        if (grid > ctx->numSMs * 8) grid = ctx->numSMs * 8;
        if (grid < 1) grid = 1;
        k_histogram<<<(unsigned)grid, 256, 0, ctx->stream>>>(d_codes, n, base, nbins, ctx->w_misc2.as<unsigned long long>(),
                                                             ctx->w_misc3.as<unsigned long long>());
        ctx->launches++;
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(counts, ctx->w_misc2.p, (size_t)nbins * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(first, ctx->w_misc3.p, (size_t)nbins * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PAC_OK;
}

// small helper for the element-wise quantiser entry points
#define ELEMWISE(IN_T, OUT_T, inptr, outptr, n, launch)                                                        \
    do {                                                                                                      \
        CK(cudaSetDevice(ctx->device));                                                                       \
        CK(ctx->w_misc.ensure((size_t)(n) * sizeof(IN_T)));                                                   \
        CK(ctx->w_misc2.ensure((size_t)(n) * sizeof(OUT_T)));                                                 \
        CK(cudaMemcpyAsync(ctx->w_misc.p, inptr, (size_t)(n) * sizeof(IN_T), cudaMemcpyHostToDevice, ctx->stream)); \
        launch;                                                                                               \
        ctx->launches++;                                                                                      \
        CK(cudaGetLastError());                                                                               \
        CK(cudaMemcpyAsync(outptr, ctx->w_misc2.p, (size_t)(n) * sizeof(OUT_T), cudaMemcpyDeviceToHost, ctx->stream)); \
        CK(cudaStreamSynchronize(ctx->stream));                                                               \
    } while (0)

extern "C" int pac_scale_factor(PacCtx *ctx, const double *x, int n, int nScaleBits, int nMantBits, int32_t *scale) {
    if (!ctx) return PAC_E_ARG;
    if (!x || !scale || n <= 0 || nScaleBits > 5 || nMantBits > 32) FAIL(PAC_E_ARG, "bad arguments to pac_scale_factor");
    ELEMWISE(double, int32_t, x, scale, n, (k_scale_factor<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->w_misc.as<double>(), n, nScaleBits, nMantBits, ctx->w_misc2.as<int32_t>())));
    return PAC_OK;
}
extern "C" int pac_vquantize_uniform(PacCtx *ctx, const double *x, int n, int nBits, uint64_t *q) {
    if (!ctx) return PAC_E_ARG;
    if (!x || !q || n <= 0 || nBits < 1 || nBits > 52) FAIL(PAC_E_ARG, "bad arguments to pac_vquantize_uniform");
    ELEMWISE(double, uint64_t, x, q, n, (k_vquantize_uniform<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->w_misc.as<double>(), n, nBits, ctx->w_misc2.as<unsigned long long>())));
    return PAC_OK;
}
extern "C" int pac_vdequantize_uniform(PacCtx *ctx, const uint64_t *q, int n, int nBits, double *x) {
    if (!ctx) return PAC_E_ARG;
    if (!x || !q || n <= 0 || nBits < 1 || nBits > 52) FAIL(PAC_E_ARG, "bad arguments to pac_vdequantize_uniform");
    ELEMWISE(uint64_t, double, q, x, n, (k_vdequantize_uniform<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->w_misc.as<unsigned long long>(), n, nBits, ctx->w_misc2.as<double>())));
    return PAC_OK;
}
extern "C" int pac_vmantissa(PacCtx *ctx, const double *x, int n, int scale, int nScaleBits, int nMantBits, uint64_t *m) {
    if (!ctx) return PAC_E_ARG;
    if (!x || !m || n <= 0 || nScaleBits > 5 || nMantBits < 1 || nMantBits > 32) FAIL(PAC_E_ARG, "bad arguments to pac_vmantissa");
    ELEMWISE(double, uint64_t, x, m, n, (k_vmantissa<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->w_misc.as<double>(), n, scale, nScaleBits, nMantBits, ctx->w_misc2.as<unsigned long long>())));
    return PAC_OK;
}
extern "C" int pac_vdequantize(PacCtx *ctx, int scale, const int64_t *m, int n, int nScaleBits, int nMantBits, double *x) {
    if (!ctx) return PAC_E_ARG;
    if (!x || !m || n <= 0 || nScaleBits > 5 || nMantBits < 1 || nMantBits > 32) FAIL(PAC_E_ARG, "bad arguments to pac_vdequantize");
    ELEMWISE(int64_t, double, m, x, n, (k_vdequantize<<<(n + 255) / 256, 256, 0, ctx->stream>>>(scale, ctx->w_misc.as<long long>(), n, nScaleBits, nMantBits, ctx->w_misc2.as<double>())));
    return PAC_OK;
}
