// decode.cuh -- K6 (chunk index + Huffman/bit unpack + dequantise) and K7 (M/S recombine + IMDCT + window +
// overlap-add + PCM quantise).
//   PACFile.ReadDataBlock pacfile.py:153-229, PackedBits.ReadBits bitpack.py:104-170, Huffman.decodeData
//   Huffman.py:321-344, codec.Decode codec.py:25-65, vDequantize quantize.py:345-376, IMDCT mdct.py:73-88,
//   SineWindow window.py:27-39, PCMFile.WriteDataBlock pcmfile.py:118-147.
#pragma once
#include "common.cuh"
#include "fft.cuh"

namespace pac {

constexpr int kLutBits = 10;
constexpr uint32_t kLutLeaf = 0x80000000u;      // entry = LEAF | (sym+1) << 8 | len   (sym -1 = escape -> 0)
constexpr uint32_t kLutInvalid = 0x7fffffffu;   // otherwise entry = trie node index reached after kLutBits bits

struct DecodeTables {
    const uint32_t *lut;        // [kNTables][1 << kLutBits]
    const int32_t *child;       // [nodes][2]   (-1 = none)
    const int32_t *sym;         // [nodes]      (-2 = internal, -1 = escape, >= 0 magnitude)
    int32_t root[kNTables];
};

// ---------------------------------------------------------------- K6a: walk the length-prefixed chunk chain
struct IndexArgs {
    const uint8_t *pac;
    const int64_t *pacBeg;      // [S] first byte of each stream's file image
    const int64_t *pacLen;      // [S] its length
    int S, hdrBytes, maxBlocks;
    int64_t *chunkPos;          // [S][maxBlocks][2] absolute byte offset of each payload
    int32_t *chunkLen;          // [S][maxBlocks][2]
    int32_t *nBlocks;           // [S]
    int32_t *status;            // [S] 0 ok, PAC_E_FORMAT on a truncated chunk
};

__device__ __forceinline__ uint32_t ld_le32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

__global__ void k_index(const IndexArgs a) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= a.S) return;
    const int64_t beg = a.pacBeg[s], end = beg + a.pacLen[s];
    int64_t pos = beg + a.hdrBytes;
    int nb = 0, st = 0;
    while (nb < a.maxBlocks) {
        int64_t p0, p1;
        uint32_t n0, n1;
        if (pos + 4 > end) break;                         // EOF on the first channel (pacfile.py:170-178)
        n0 = ld_le32(a.pac + pos); p0 = pos + 4; pos = p0 + n0;
        if (pos > end) { st = PAC_E_FORMAT; break; }       // pacfile.py:184
        if (pos + 4 > end) break;                         // EOF on the second channel: block is dropped
        n1 = ld_le32(a.pac + pos); p1 = pos + 4; pos = p1 + n1;
        if (pos > end) { st = PAC_E_FORMAT; break; }
        int64_t o = ((int64_t)s * a.maxBlocks + nb) * 2;
        a.chunkPos[o] = p0; a.chunkPos[o + 1] = p1;
        a.chunkLen[o] = (int32_t)n0; a.chunkLen[o + 1] = (int32_t)n1;
        nb++;
    }
    a.nBlocks[s] = nb;
    a.status[s] = st;
}

// gather every stream's file header into one contiguous buffer (one D2H instead of S)
__global__ void k_gather_headers(const uint8_t *pac, const int64_t *pacBeg, int S, int hdrBytes, uint8_t *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)S * hdrBytes) return;
    int s = (int)(i / hdrBytes), j = (int)(i - (int64_t)s * hdrBytes);
    out[i] = pac[pacBeg[s] + j];
}

// ---------------------------------------------------------------- K6b: one thread parses one channel chunk
struct BitReader {
    const uint8_t *p;
    int64_t nbits, pos;
    __device__ __forceinline__ uint32_t peek(int n) {      // next n (<= 24) bits, zero padded past the end
        int64_t byte = pos >> 3;
        int sh = (int)(pos & 7);
        uint64_t v = 0;
#pragma unroll
        for (int i = 0; i < 5; i++) {
            int64_t bi = byte + i;
            uint64_t bv = (bi * 8 < nbits) ? p[bi] : 0;
            v = (v << 8) | bv;
        }
        return (uint32_t)((v >> (40 - sh - n)) & ((1u << n) - 1u));
    }
    __device__ __forceinline__ uint32_t get(int n) { uint32_t v = n ? peek(n) : 0; pos += n; return v; }
};

template <typename T>
struct UnpackArgs {
    const uint8_t *pac;
    const int64_t *chunkPos;     // [nchunk] absolute offsets
    const int32_t *chunkLen;     // [nchunk]
    const int32_t *nBlocks;      // [S] (NULL: all chunks valid)
    int S, maxBlocks, M;
    int nScaleBits, nMantSizeBits, nTableIDBits;
    // outputs, chunk-indexed c = (s*maxBlocks + b)*2 + ch
    T *lines;                    // [nchunk][M] dequantised, / 2^overallScale (may be NULL)
    uint32_t *lrms;              // [nblocks] (from the LAST channel parsed, pacfile.py:216-217)
    int32_t *err;                // [S]
    int32_t *o_sf, *o_ba, *o_mant, *o_oscale, *o_tableID;   // optional raw fields (per-block API)
    DecodeTables dt;
    BandInfo bands;
};

template <typename T>
__global__ void k_unpack(const UnpackArgs<T> a) {
    const int64_t nchunk = (int64_t)a.S * a.maxBlocks * 2;
    const int NB = a.bands.nBands, M = a.M;
    const int largestScale = (1 << a.nScaleBits) - 1;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunk; c += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = c >> 1;
        const int s = (int)(w / a.maxBlocks);
        const int b = (int)(w - (int64_t)s * a.maxBlocks);
        if (a.nBlocks && b >= a.nBlocks[s]) continue;
        BitReader r;
        r.p = a.pac + a.chunkPos[c];
        r.nbits = (int64_t)a.chunkLen[c] * 8;
        r.pos = 0;
        const int oscale = (int)r.get(a.nScaleBits);                  // pacfile.py:187
        const int tid = (int)r.get(a.nTableIDBits);                   // :190
        bool bad = tid < 1 || tid > kNTables;
        const uint32_t *lut = a.dt.lut + (size_t)(bad ? 0 : tid - 1) * (1 << kLutBits);
        const double rescale = 1.0 / (double)(1 << oscale);           // codec.py:32,43 (exact power of two)
        T *out = a.lines ? a.lines + c * M : nullptr;
        if (a.o_oscale) { a.o_oscale[c] = oscale; a.o_tableID[c] = tid; }
        for (int bd = 0; bd < NB && !bad; bd++) {
            int ba = (int)r.get(a.nMantSizeBits);                     // :195
            if (ba) ba += 1;                                          // :196
            const int sf = (int)r.get(a.nScaleBits);                  // :198
            if (a.o_ba) { a.o_ba[c * kMaxBands + bd] = ba; a.o_sf[c * kMaxBands + bd] = sf; }
            const int lo = a.bands.lo[bd], hi = a.bands.lo[bd + 1];
            if (!ba) {
                for (int i = lo; i < hi; i++) { if (out) out[i] = (T)0; if (a.o_mant) a.o_mant[c * M + i] = 0; }
                continue;
            }
            const int64_t signPos = r.pos;                            // nLines sign bits first (:202-204)
            r.pos += hi - lo;
            for (int i = lo; i < hi; i++) {
                uint32_t e = lut[r.peek(kLutBits)];
                int sym;
                if (e & kLutLeaf) { r.pos += e & 0xff; sym = (int)((e >> 8) & 0x7fffff) - 1; }
                else if (e == kLutInvalid) { bad = true; break; }
                else {                                                // long code: continue bit-serially (Huffman.py:337-344)
                    int node = (int)e;
                    r.pos += kLutBits;
                    while (a.dt.sym[node] == -2) {
                        node = a.dt.child[2 * node + (int)r.get(1)];
                        if (node < 0 || r.pos > r.nbits) { bad = true; break; }
                    }
                    if (bad) break;
                    sym = a.dt.sym[node];
                }
                long long m = sym < 0 ? (long long)r.get(ba) : (long long)sym;    // escape: Huffman.py:326-327
                BitReader sr = r;
                sr.pos = signPos + (i - lo);
                if (sr.peek(1)) m += 1ll << (ba - 1);                 // pacfile.py:210
                if (a.o_mant) a.o_mant[c * M + i] = (int32_t)m;
                if (out) out[i] = (T)(dequant(sf, m, largestScale, ba) * rescale);
            }
        }
        uint32_t lr = 0;
        for (int bd = 0; bd < NB; bd++) lr |= r.get(1) << bd;         // :216-217
        if (r.pos > r.nbits) bad = true;
        if ((c & 1) == 1) a.lrms[w] = lr;                             // the last channel's copy wins
        if (bad && a.err) a.err[s] = PAC_E_FORMAT;
    }
}

// ---------------------------------------------------------------- K7: synthesis
template <typename T>
struct SynthArgs {
    const T *lines;              // [S][maxBlocks][2][M] dequantised (pre M/S recombination)
    const uint32_t *lrms;        // [S][maxBlocks]
    const int32_t *nBlocks;      // [S]
    int S, maxBlocks, run;       // each CTA produces `run` consecutive output blocks of one stream
    int16_t *pcm;                // [S][strideSamples][2]
    int64_t strideSamples;
    int64_t *nSamplesOut;        // [S]
    double *rawOut;              // per-block API: [nblk][2][N] windowed IMDCT output, no overlap-add
    DevTables<T> tab;
    BandInfo bands;
};

template <typename T, int LOGM>
struct SynthSmem {
    static constexpr int M = 1 << LOGM;
    using T2 = typename Vec2<T>::type;
    T X[2][M];
    T2 W[2][M / 2 + 2];
    T y[2][2 * M];
    T ola[2][M];
};

// PCMFile.WriteDataBlock quantisation (pcmfile.py:127-134, quantize.py:91-117 with 16 bits)
__device__ __forceinline__ int pcm16(double v) {
    double a = fabs(v);
    int code = a < 1.0 ? (int)((a * 65535.0 + 1.0) / 2.0) : 32767;
    return signbit(v) ? -code : code;
}

template <typename T, int LOGM>
__global__ void __launch_bounds__((1 << LOGM) / 4)
k_synth(const SynthArgs<T> a) {
    using SS = SynthSmem<T, LOGM>;
    using T2 = typename Vec2<T>::type;
    constexpr int M = SS::M, N = 2 * M, NT = M / 4, H = M / 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SS &sm = *reinterpret_cast<SS *>(smem_raw);
    const int tid = threadIdx.x;
    const DevTables<T> &tb = a.tab;
    const int runsPerStream = a.rawOut ? a.maxBlocks : (a.maxBlocks + a.run) / a.run;   // +1 output block for the tail
    const int s = blockIdx.x / runsPerStream;
    const int r = blockIdx.x - s * runsPerStream;
    if (s >= a.S) return;
    const int nblk = a.nBlocks ? a.nBlocks[s] : a.maxBlocks;
    // Output block j (j = 0..nblk-1) = ola(block j) + first half(block j+1)  for j < nblk-1  [block 0's own
    // output is dropped, pacfile.py:485-487], and the last output block is the saved tail (pacfile.py:171-176).
    // In block terms: out[j] = second_half(dec[j]) + first_half(dec[j+1]), j = 0..nblk-2; out[nblk-1] = second_half(dec[nblk-1]).
    int j0, j1;
    if (a.rawOut) { j0 = r; j1 = r + 1; }
    else { j0 = r * a.run; j1 = min(j0 + a.run, nblk); if (j0 >= nblk) { return; } }
    if (!a.rawOut && r == 0 && tid == 0) a.nSamplesOut[s] = (int64_t)nblk * M;
    const int bFirst = j0, bLast = a.rawOut ? j0 : min(j1, nblk - 1);    // decoded blocks needed: j0 .. j1 (clipped)
    for (int b = bFirst; b <= bLast; b++) {
        const int64_t w = (int64_t)s * a.maxBlocks + b;
        const uint32_t lrms = a.lrms[w];
        // load + M/S recombination with the reference's aliasing (codec.py:46-56): L' = M - S, R' = L' + S
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int i = tid + NT * j;
            T x0 = a.lines[(w * 2 + 0) * M + i], x1 = a.lines[(w * 2 + 1) * M + i];
            if ((lrms >> tb.band_of_line[i]) & 1u) { x0 = x0 - x1; x1 = x0 + x1; }
            sm.X[0][i] = x0; sm.X[1][i] = x1;
        }
        __syncthreads();
        // DCT-IV by fold + M/2-point FFT (same routine as the forward MDCT)
        for (int e = tid; e < 2 * H; e += NT) {
            int ch = e / H, n = e - ch * H;
            sm.W[ch][n] = cmul(mk2<T>(sm.X[ch][2 * n], sm.X[ch][M - 1 - 2 * n]), tb.mdct_pre[n]);
        }
        __syncthreads();
        fft_dif<T, LOGM - 1, NT>(&sm.W[0][0], 2, H + 2, tb.tw, 2);
        for (int e = tid; e < 2 * H; e += NT) {
            int ch = e / H, k = e - ch * H;
            T2 yv = cmul(sm.W[ch][fft_pos<LOGM - 1>(k)], tb.mdct_post[k]);
            sm.X[ch][2 * k] = yv.x;                 // v[2k]
            sm.X[ch][M - 1 - 2 * k] = -yv.y;        // v[M-1-2k]
        }
        __syncthreads();
        // unfold (IMDCT, mdct.py:73-80: y[n] = 2 sum_k X[k] cos(2pi/N (n+n0)(k+1/2))) and SineWindow (codec.py:59-60)
        for (int e = tid; e < 2 * N; e += NT) {
            int ch = e / N, n = e - ch * N;
            T v = n < H ? sm.X[ch][n + H] : (n < 3 * H ? -sm.X[ch][3 * H - 1 - n] : -sm.X[ch][n - 3 * H]);
            sm.y[ch][n] = (T)2 * v * tb.sinw[n];
        }
        __syncthreads();
        if (a.rawOut) {
            for (int e = tid; e < 2 * N; e += NT) a.rawOut[(int64_t)blockIdx.x * 2 * N + e] = (double)sm.y[e / N][e % N];
            return;
        }
        // overlap-add (pacfile.py:223-226) and PCM quantisation; output block index = b - 1
        if (b > j0) {
            int16_t *dst = a.pcm + ((int64_t)s * a.strideSamples + (int64_t)(b - 1) * M) * 2;
            for (int e = tid; e < 2 * M; e += NT) {
                int i = e >> 1, ch = e & 1;
                dst[e] = (int16_t)pcm16((double)sm.ola[ch][i] + (double)sm.y[ch][i]);
            }
        }
        __syncthreads();
        for (int e = tid; e < 2 * M; e += NT) sm.ola[e / M][e % M] = sm.y[e / M][M + e % M];
        __syncthreads();
    }
    // the tail block (EOF): only the CTA whose range reaches the end emits it
    if (j1 == nblk) {
        int16_t *dst = a.pcm + ((int64_t)s * a.strideSamples + (int64_t)(nblk - 1) * M) * 2;
        for (int e = tid; e < 2 * M; e += NT) dst[e] = (int16_t)pcm16((double)sm.ola[e & 1][e >> 1]);
    }
}

}  // namespace pac
