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
    int64_t *chunkPos;          // [S][maxBlocks][2] absolute byte offset of each payload (NULL: count blocks only)
    int32_t *chunkLen;          // [S][maxBlocks][2]
    int32_t *nBlocks;           // [S]
    int32_t *status;            // [S] 0 ok, PAC_E_FORMAT on a truncated chunk
};

__device__ __forceinline__ uint32_t ld_le32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

// One warp per stream.  The chain is serial (each length prefix tells where the next one is), so the warp stages a
// 4 KB window of the image in shared memory with coalesced 16 B loads and every lane walks it redundantly from
// there (uniform control flow); a new window is fetched only when the walk leaves the current one (~13 links).
constexpr int kIdxWin = 4096, kIdxWarps = 4;
constexpr int kIdxBoundExceeded = -100;          // internal status: the chain holds more blocks than IndexArgs.maxBlocks

__global__ void __launch_bounds__(32 * kIdxWarps) k_index(const IndexArgs a) {
    __shared__ __align__(16) uint8_t win[kIdxWarps][kIdxWin];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * kIdxWarps + warp;
    if (s >= a.S) return;
    const int64_t beg = a.pacBeg[s], end = beg + a.pacLen[s];
    const uintptr_t base = (uintptr_t)a.pac;
    const uintptr_t aEnd = (base + (uintptr_t)end + 15) & ~(uintptr_t)15;   // loads stay inside the image's 16 B granules
    uintptr_t wlo = 0, whi = 0;                                             // window = addresses [wlo, whi)
    uint8_t *w = win[warp];
    auto rd = [&](int64_t p) -> uint32_t {                                  // little-endian u32 at image offset p (p+4 <= end)
        const uintptr_t A = base + (uintptr_t)p;
        if (A < wlo || A + 4 > whi) {
            __syncwarp();
            wlo = A & ~(uintptr_t)15; whi = wlo + kIdxWin;
#pragma unroll
            for (int k = 0; k < kIdxWin / 512; k++) {
                uintptr_t q = wlo + (uintptr_t)(k * 512 + lane * 16);
                if (q < aEnd) *reinterpret_cast<uint4 *>(w + k * 512 + lane * 16) = __ldg(reinterpret_cast<const uint4 *>(q));
            }
            __syncwarp();
        }
        return ld_le32(w + (A - wlo));
    };
    int64_t pos = beg + a.hdrBytes;
    int nb = 0, st = 0;
    while (nb < a.maxBlocks) {
        int64_t p0, p1;
        uint32_t n0, n1;
        if (pos + 4 > end) break;                         // EOF on the first channel (pacfile.py:170-178)
        n0 = rd(pos); p0 = pos + 4; pos = p0 + n0;
        if (pos > end) { st = PAC_E_FORMAT; break; }       // pacfile.py:184
        if (pos + 4 > end) break;                         // EOF on the second channel: block is dropped
        n1 = rd(pos); p1 = pos + 4; pos = p1 + n1;
        if (pos > end) { st = PAC_E_FORMAT; break; }
        if (lane == 0 && a.chunkPos) {
            int64_t o = ((int64_t)s * a.maxBlocks + nb) * 2;
            *reinterpret_cast<longlong2 *>(a.chunkPos + o) = make_longlong2(p0, p1);
            *reinterpret_cast<int2 *>(a.chunkLen + o) = make_int2((int)n0, (int)n1);
        }
        nb++;
    }
    // more chunks than the caller's bound allows: never silently truncated.  The bound is either generous (from the image length: then
    // the file is malformed) or a hint from the header's sample count (then the caller walks again, counting first)
    if (st == 0 && nb == a.maxBlocks && pos + 4 <= end) {
        const uint32_t n0 = rd(pos);
        if (pos + 4 + (int64_t)n0 + 4 <= end) st = kIdxBoundExceeded;
    }
    if (lane == 0) { a.nBlocks[s] = nb; a.status[s] = st; }
}

// gather every stream's file header into one contiguous buffer (one D2H instead of S)
__global__ void k_gather_headers(const uint8_t *pac, const int64_t *pacBeg, int S, int hdrBytes, uint8_t *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)S * hdrBytes) return;
    int s = (int)(i / hdrBytes), j = (int)(i - (int64_t)s * hdrBytes);
    out[i] = pac[pacBeg[s] + j];
}

// ---------------------------------------------------------------- K6b: one thread parses one channel chunk
// MSB-first bit reader (bitpack.py:104-170) over global memory: the 32-bit word holding the current bit, the next
// one and a prefetched third (its load latency hides behind the 32 bits in front of it); aligned, clamped loads.
// Reading past the chunk returns repeated bytes; `used > nbits` flags it afterwards.
struct BitReader {
    const uint32_t *wp, *wlast, *wbase;
    uint32_t w0, w1, w2;
    int bp;                      // bit offset of the cursor inside w0
    int off0, nbits;             // bit offset of the chunk's first bit inside *wbase; chunk size in bits
    __device__ __forceinline__ uint32_t word() {
        uint32_t v = __ldg(wp < wlast ? wp : wlast);
        wp++;
        return __byte_perm(v, 0, 0x0123);
    }
    __device__ __forceinline__ void seek(const uint8_t *p, int bitpos) {
        uintptr_t A = (uintptr_t)p + (uintptr_t)(bitpos >> 3);
        wp = reinterpret_cast<const uint32_t *>(A & ~(uintptr_t)3);
        bp = (int)(A & 3) * 8 + (bitpos & 7);
        w0 = word(); w1 = word(); w2 = word();
    }
    __device__ __forceinline__ void init(const uint8_t *p, int nbytes) {
        nbits = nbytes * 8;
        wlast = reinterpret_cast<const uint32_t *>(((uintptr_t)p + (uintptr_t)(nbytes > 0 ? nbytes - 1 : 0)) & ~(uintptr_t)3);
        wbase = reinterpret_cast<const uint32_t *>((uintptr_t)p & ~(uintptr_t)3);
        off0 = (int)((uintptr_t)p & 3) * 8;
        seek(p, 0);
    }
    // bits consumed since the start of the chunk, from the cursor itself (w0 is the word three behind wp): nothing to maintain per token
    __device__ __forceinline__ int used() const { return ((int)(wp - wbase) - 3) * 32 + bp - off0; }
    __device__ __forceinline__ uint32_t peek(int n) { return __funnelshift_l(w1, w0, bp) >> (32 - n); }     // 1 <= n <= 32
    __device__ __forceinline__ void skip(int n) {                                                           // 0 <= n <= 32
        bp += n;
        if (bp >= 32) { bp -= 32; w0 = w1; w1 = w2; w2 = word(); }
    }
    __device__ __forceinline__ uint32_t get(int n) {
        if (!n) return 0;
        uint32_t v = peek(n);
        skip(n);
        return v;
    }
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
    uint16_t *codes;             // [nchunk][M] mantissa codes, sign at bit ba-1 (lines of zero-bit bands may be left unwritten)
    uint16_t *meta;              // [nchunk][kMaxBands] sf | ba << 8
    uint8_t *oscale;             // [nchunk]
    uint32_t *lrms;              // [nblocks] (from the LAST channel parsed, pacfile.py:216-217)
    int32_t *err;                // [S]
    int32_t *o_sf, *o_ba, *o_mant, *o_oscale, *o_tableID;   // optional raw fields (per-block API)
    DecodeTables dt;
    BandInfo bands;
};

// Collects a chunk's codes in line order and writes them 8 at a time (one 16 B store): a thread owns a whole 2 KB row,
// so 2-byte stores would cost one L2 sector write each.
struct CodeSink {
    uint16_t *row;
    uint4 q;
    uint32_t cur;
    __device__ __forceinline__ void push(uint32_t v, int i) {
        if (i & 1) {
            q.x = q.y; q.y = q.z; q.z = q.w; q.w = cur | (v << 16);
            if ((i & 7) == 7) *reinterpret_cast<uint4 *>(row + i - 7) = q;
        } else cur = v;
    }
};

constexpr int kUnpackThreads = 128;

// (The first-level tables stay in global memory: a shared-memory copy (40 KB) cost more in occupancy and L1 capacity
// than it saved in lookup latency -- 52.6 vs 26.7 ms on the 1024 x 60 s corpus.)
template <typename T>
__global__ void __launch_bounds__(kUnpackThreads) k_unpack(const UnpackArgs<T> a) {
    const int64_t nchunk = (int64_t)a.S * a.maxBlocks * 2;
    const int NB = a.bands.nBands, M = a.M;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunk; c += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = c >> 1;
        const int s = (int)(w / a.maxBlocks);
        const int b = (int)(w - (int64_t)s * a.maxBlocks);
        if (a.nBlocks && b >= a.nBlocks[s]) continue;
        const uint8_t *p = a.pac + a.chunkPos[c];
        BitReader r;
        r.init(p, a.chunkLen[c]);
        bool bad = a.chunkLen[c] <= 0;
        const int oscale = (int)r.get(a.nScaleBits);                  // pacfile.py:187
        const int tid = (int)r.get(a.nTableIDBits);                   // :190
        bad |= tid < 1 || tid > kNTables;
        const uint32_t *lut = a.dt.lut + (size_t)(bad ? 0 : tid - 1) * (1 << kLutBits);
        CodeSink sink;
        sink.row = a.codes ? a.codes + c * M : nullptr;
        sink.q = make_uint4(0, 0, 0, 0); sink.cur = 0;
        if (a.oscale) a.oscale[c] = (uint8_t)oscale;
        if (a.o_oscale) { a.o_oscale[c] = oscale; a.o_tableID[c] = tid; }
        for (int bd = 0; bd < NB && !bad; bd++) {
            int ba = (int)r.get(a.nMantSizeBits);                     // :195
            if (ba) ba += 1;                                          // :196
            const int sf = (int)r.get(a.nScaleBits);                  // :198
            if (a.meta) a.meta[c * kMaxBands + bd] = (uint16_t)(sf | (ba << 8));
            if (a.o_ba) { a.o_ba[c * kMaxBands + bd] = ba; a.o_sf[c * kMaxBands + bd] = sf; }
            const int lo = a.bands.lo[bd], hi = a.bands.lo[bd + 1];
            if (!ba) {
                if (a.o_mant) for (int i = lo; i < hi; i++) a.o_mant[c * M + i] = 0;
                if (sink.row) {                                       // keep the 8-line groups that straddle the band edges whole;
                    int i = lo;                                       // groups entirely inside a zero-bit band are never read
                    for (; i < hi && (i & 7); i++) sink.push(0, i);
                    i += (hi - i) & ~7;
                    for (; i < hi; i++) sink.push(0, i);
                }
                continue;
            }
            BitReader sr = r;                                         // nLines sign bits first (:202-204) ...
            r.seek(p, r.used() + (hi - lo));                            // ... then the codes
            const uint32_t signBit = 1u << (ba - 1);
            uint32_t sbits = 0;                                       // up to 32 buffered sign bits, next one in bit 31
            int sleft = 0;
            int i = lo;
            while (i < hi) {
                if (sleft == 0) { const int n = min(32, hi - i); sbits = sr.get(n) << (32 - n); sleft = n; }
                uint32_t e = lut[r.peek(kLutBits)];
                int sym;
                if (e & kLutLeaf) { r.skip((int)(e & 0xff)); sym = (int)((e >> 8) & 0x7fffff) - 1; }
                else if (e == kLutInvalid) { bad = true; break; }
                else {                                                // long code: continue bit-serially (Huffman.py:337-344)
                    int node = (int)e;
                    r.skip(kLutBits);
                    while (a.dt.sym[node] == -2) {
                        node = a.dt.child[2 * node + (int)r.get(1)];
                        if (node < 0 || r.used() > r.nbits) { bad = true; break; }
                    }
                    if (bad) break;
                    sym = a.dt.sym[node];
                }
                uint32_t m = (uint32_t)sym;
                if (sym < 0) {                                        // escape: the magnitude follows in ba raw bits (Huffman.py:326-327)
                    m = r.get(ba);
                    if (m >= signBit) { bad = true; break; }          // its top bit is always 0 (Huffman.py:296-298 writes ba bits of a
                }                                                     // ba-1 bit magnitude); table symbols are < 2^15 by construction
                if (sbits >> 31) m += signBit;                        // pacfile.py:210
                sbits <<= 1; sleft--;
                if (a.o_mant) a.o_mant[c * M + i] = (int32_t)m;
                if (sink.row) sink.push(m, i);
                i++;
            }
            if (bad) break;
        }
        uint32_t lr = 0;
        for (int bd = 0; bd < NB; bd++) lr |= r.get(1) << bd;         // :216-217
        if (r.used() > r.nbits) bad = true;
        if ((c & 1) == 1) a.lrms[w] = lr;                             // the last channel's copy wins
        if (bad && a.err) a.err[s] = PAC_E_FORMAT;
    }
}

// ---------------------------------------------------------------- K7: synthesis
template <typename T>
struct SynthArgs {
    const T *lines;              // [S][maxBlocks][2][M] dequantised (pre M/S recombination), or NULL and instead:
    const uint16_t *codes;       // [chunk][M]  k_unpack's mantissa codes
    const uint16_t *meta;        // [chunk][kMaxBands] sf | ba << 8
    const uint8_t *oscale;       // [chunk]
    int largestScale;            // (1 << nScaleBits) - 1
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
    // (fp32, 1024 lines: the register-blocked radix-8 path addresses W and v through XOR swizzles, k_synth below; the spare elements
    // of the rows are left over from the additive paddings it used before and keep the rows' alignment)
    static constexpr int WROW = M / 2 + M / 16 + 2;
    T2 W[2][WROW];               // folded spectrum / FFT workspace
    T v[2][M + M / 16];          // DCT-IV output
    uint16_t meta[2][kMaxBands];
    float gain[2][kMaxBands];    // fp32 mode: 2/(2^R-1) * 2^(largestScale-sf-1-overallScale) per band
    // double-buffered cp.async landing area: block b+1's codes and band metadata arrive while block b is transformed
    __align__(16) uint16_t codes[2][2][M];
    __align__(16) uint16_t metaRaw[2][2 * kMaxBands];
};

__device__ __forceinline__ void cp_async16(void *smemDst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smemDst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// PCMFile.WriteDataBlock quantisation (pcmfile.py:127-134, quantize.py:91-117 with 16 bits)
__device__ __forceinline__ int pcm16(double v) {
    double a = fabs(v);
    int code = a < 1.0 ? (int)((a * 65535.0 + 1.0) / 2.0) : 32767;
    return signbit(v) ? -code : code;
}
__device__ __forceinline__ int pcm16(float v) {          // fp32 mode: within 1 LSB of the above
    float a = fabsf(v);
    int code = a < 1.0f ? (int)(a * 32767.5f + 0.5f) : 32767;
    return signbit(v) ? -code : code;
}

// one dequantised line (codec.py:31-43, quantize.py:345-376)
template <typename T, typename SS>
__device__ __forceinline__ T synth_line(const SynthArgs<T> &a, const SS &sm, int64_t chunk, int ch, int i, int bd, double rescale, int buf) {
    if (!a.codes) return a.lines[chunk * SS::M + i];
    const uint32_t mt = sm.meta[ch][bd];
    const int ba = (int)(mt >> 8), sf = (int)(mt & 0xff);
    if (!ba) return (T)0;
    const uint32_t code = sm.codes[buf][ch][i];
    if constexpr (sizeof(T) == 8) {
        return (T)(dequant(sf, (long long)code, a.largestScale, ba) * rescale);
    } else {
        const uint32_t signMask = 1u << (ba - 1);
        const bool neg = (code & signMask) != 0;
        const uint32_t mag = neg ? code - signMask : code;
        // q = mag << (L - sf) (+ half a step)  =  (2 mag + 1) * 2^(L-sf-1), exact in fp32; the power of two lives in gain[]
        const float q = (float)(int)(2u * mag + ((sf < a.largestScale && mag > 0) ? 1u : 0u));
        const float x = q * sm.gain[ch][bd];
        return neg ? -x : x;
    }
}

template <typename T, int LOGM>
__global__ void __launch_bounds__((1 << LOGM) / 4, sizeof(T) == 4 ? 5 : 1)
k_synth(const SynthArgs<T> a) {
    using SS = SynthSmem<T, LOGM>;
    using T2 = typename Vec2<T>::type;
    constexpr int M = SS::M, N = 2 * M, NT = M / 4, H = M / 2;
    constexpr bool R8 = sizeof(T) == 4 && LOGM == 10;        // fp32, 1024 lines: register-blocked radix-8 FFT (see below)
    // Shared-memory layouts of the register-blocked path: XOR swizzles found by enumerating the access patterns (every load and store below
    // then takes the minimum number of wavefronts).  v (floats): the last FFT pass writes v[2k], v[M-1-2k] for k = 8(t&7) + (t>>3) + 64p,
    // the unfold reads ascending/descending runs; W (float2, half-warp per wavefront): linear n, t + 64r, 64g + j + 8r, 8t + r.
    // (The additive paddings used before -- one spare per 16 / per 8 -- left the v stores 4-way and the linear W accesses 2-way
    // conflicted: 39 % of the kernel's shared-memory wavefronts were replays.)
    auto VI = [](int i) { return R8 ? i ^ ((i >> 2) & 31) : i; };   // position of v[i]
    auto ZI = [](int i) { return i ^ ((i >> 3) & 15); };             // position of W[ch][i] (R8 only)
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
    T ola[4][2];                                                         // overlap tail of samples tid + NT*j (this thread's own)
#pragma unroll
    for (int j = 0; j < 4; j++) ola[j][0] = ola[j][1] = (T)0;
    // block b+1's codes (4 KB: one 16-byte cp.async per thread) and band metadata travel while block b is transformed; its
    // LRMS mask and overall scales are ordinary loads issued one iteration ahead
    auto prefetch = [&](int b, int buf) {
        const int64_t w = (int64_t)s * a.maxBlocks + b;
        cp_async16(&sm.codes[buf][0][0] + 8 * tid, a.codes + w * 2 * M + 8 * tid);
        if (tid < (2 * kMaxBands * 2) / 16) cp_async16(&sm.metaRaw[buf][0] + 8 * tid, a.meta + w * 2 * kMaxBands + 8 * tid);
        cp_async_commit();
    };
    uint32_t lrmsNext = 0, oscNext = 0;
    if (bFirst <= bLast) {
        const int64_t w = (int64_t)s * a.maxBlocks + bFirst;
        if (a.codes) { prefetch(bFirst, 0); oscNext = (uint32_t)a.oscale[w * 2] | (uint32_t)a.oscale[w * 2 + 1] << 8; }
        lrmsNext = a.lrms[w];
    }
    for (int b = bFirst; b <= bLast; b++) {
        const int64_t w = (int64_t)s * a.maxBlocks + b;
        const uint32_t lrms = lrmsNext;
        const int osc0 = (int)(oscNext & 0xffu), osc1 = (int)(oscNext >> 8);
        const int buf = (b - bFirst) & 1;
        double r0 = 1.0, r1 = 1.0;
        if (a.codes) {
            cp_async_wait_all();
            __syncthreads();                       // this block's codes/metadata have landed; the other buffer's readers are done
        }
        if (b < bLast) {
            if (a.codes) { prefetch(b + 1, buf ^ 1); oscNext = (uint32_t)a.oscale[(w + 1) * 2] | (uint32_t)a.oscale[(w + 1) * 2 + 1] << 8; }
            lrmsNext = a.lrms[w + 1];
        }
        if (a.codes) {
            if (tid < 2 * kMaxBands) {
                const int ch = tid / kMaxBands, bd = tid % kMaxBands;
                const uint32_t mt = sm.metaRaw[buf][tid];
                sm.meta[ch][bd] = (uint16_t)mt;
                if (sizeof(T) == 4) {
                    const int ba = (int)(mt >> 8), sf = (int)(mt & 0xff);
                    const double largest = (double)(1ll << ((ba + a.largestScale) & 63)) - 1.0;
                    sm.gain[ch][bd] = (float)ldexp(2.0 / largest, a.largestScale - sf - 1 - (ch ? osc1 : osc0));
                }
            }
            r0 = 1.0 / (double)(1 << osc0); r1 = 1.0 / (double)(1 << osc1);   // exact powers of two
            __syncthreads();
        }
        // load (+ dequantise) + M/S recombination with the reference's aliasing (codec.py:46-56: L' = M - S,
        // R' = L' + S), folded straight into the DCT-IV pre-twiddle: W[n] = (X[2n] + i X[M-1-2n]) e^{-i pi n/M}
        for (int n = tid; n < H; n += NT) {
            const int iA = 2 * n, iB = M - 1 - 2 * n;
            const int bA = tb.band_of_line[iA], bB = tb.band_of_line[iB];
            T a0 = synth_line<T, SS>(a, sm, w * 2, 0, iA, bA, r0, buf), a1 = synth_line<T, SS>(a, sm, w * 2 + 1, 1, iA, bA, r1, buf);
            T b0 = synth_line<T, SS>(a, sm, w * 2, 0, iB, bB, r0, buf), b1 = synth_line<T, SS>(a, sm, w * 2 + 1, 1, iB, bB, r1, buf);
            if ((lrms >> bA) & 1u) { a0 = a0 - a1; a1 = a0 + a1; }
            if ((lrms >> bB) & 1u) { b0 = b0 - b1; b1 = b0 + b1; }
            const T2 pre = tb.mdct_pre[n];
            sm.W[0][R8 ? ZI(n) : n] = cmul(mk2<T>(a0, b0), pre);
            sm.W[1][R8 ? ZI(n) : n] = cmul(mk2<T>(a1, b1), pre);
        }
        __syncthreads();
        if constexpr (R8) {
            // H = 512 complex points per channel as radix 8 x 8 x 8, one 8-point DFT per thread and pass held in registers (threads
            // 0..63 channel 0, 64..127 channel 1; the other half of the CTA idles through the three short passes): 3 passes through
            // shared memory instead of 5, and the last one applies the DCT-IV post-twiddle and writes v directly.
            if (tid < 2 * (H / 8)) {
                const int ch = tid / (H / 8), t = tid - ch * (H / 8);
                T2 *Z = sm.W[ch];
                T2 x[8];
#pragma unroll
                for (int r = 0; r < 8; r++) x[r] = Z[ZI(t + (H / 8) * r)];
                dft8(x);
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    T2 y = x[brev3(p)];
                    if (p) y = cmul(y, tb.tw[2 * t * p]);                 // W_H^(t p), tw[m] = exp(-2 pi i m / M)
                    Z[ZI(t + (H / 8) * p)] = y;
                }
            }
            __syncthreads();
            if (tid < 2 * (H / 8)) {
                const int ch = tid / (H / 8), t = tid - ch * (H / 8);
                const int g = t >> 3, j = t & 7;
                T2 *Z = sm.W[ch];
                T2 x[8];
#pragma unroll
                for (int r = 0; r < 8; r++) x[r] = Z[ZI(64 * g + j + 8 * r)];
                dft8(x);
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    T2 y = x[brev3(p)];
                    if (p) y = cmul(y, tb.tw[16 * j * p]);                // W_64^(j p)
                    Z[ZI(64 * g + j + 8 * p)] = y;
                }
            }
            __syncthreads();
            if (tid < 2 * (H / 8)) {
                const int ch = tid / (H / 8), t = tid - ch * (H / 8);
                const T2 *Z = sm.W[ch];
                T2 x[8];
#pragma unroll
                for (int r = 0; r < 8; r++) x[r] = Z[ZI(8 * t + r)];
                dft8(x);
                const int kb = (t >> 3) + 8 * (t & 7);                    // frequency k = kb + 64 p
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    const int k = kb + 64 * p;
                    const T2 yv = cmul(x[brev3(p)], tb.mdct_post[k]);
                    sm.v[ch][VI(2 * k)] = yv.x;                           // v[2k]
                    sm.v[ch][VI(M - 1 - 2 * k)] = -yv.y;                  // v[M-1-2k]
                }
            }
        } else {
            fft_dif<T, LOGM - 1, NT>(&sm.W[0][0], 2, SS::WROW, tb.tw, 2);
            for (int e = tid; e < 2 * H; e += NT) {
                int ch = e / H, k = e - ch * H;
                T2 yv = cmul(sm.W[ch][fft_pos<LOGM - 1>(k)], tb.mdct_post[k]);
                sm.v[ch][2 * k] = yv.x;                 // v[2k]
                sm.v[ch][M - 1 - 2 * k] = -yv.y;        // v[M-1-2k]
            }
        }
        __syncthreads();
        // unfold (IMDCT, mdct.py:73-80: y[n] = 2 sum_k X[k] cos(2pi/N (n+n0)(k+1/2))), SineWindow (codec.py:59-60),
        // overlap-add (pacfile.py:223-226) and PCM quantisation in one pass; output block index = b - 1
        int16_t *dst = a.rawOut ? nullptr : a.pcm + ((int64_t)s * a.strideSamples + (int64_t)(b - 1) * M) * 2;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = tid + NT * j;
            const int iF = i < H ? i + H : 3 * H - 1 - i;      // y[i]   =  v[i+H] | -v[3H-1-i]
            const int iS = i < H ? H - 1 - i : i - H;          // y[M+i] = -v[H-1-i] | -v[i-H]
            const T wF = tb.sinw[i], wS = tb.sinw[M + i];
            short2 o;
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                T vF = sm.v[ch][VI(iF)], vS = -sm.v[ch][VI(iS)];
                if (i >= H) vF = -vF;
                const T yF = (T)2 * vF * wF, yS = (T)2 * vS * wS;
                if (a.rawOut) {
                    a.rawOut[((int64_t)blockIdx.x * 2 + ch) * N + i] = (double)yF;
                    a.rawOut[((int64_t)blockIdx.x * 2 + ch) * N + M + i] = (double)yS;
                } else {
                    int q;
                    if constexpr (sizeof(T) == 8) q = pcm16((double)ola[j][ch] + (double)yF);
                    else q = pcm16(ola[j][ch] + yF);
                    if (ch == 0) o.x = (short)q; else o.y = (short)q;
                    ola[j][ch] = yS;
                }
            }
            if (!a.rawOut && b > j0) *reinterpret_cast<short2 *>(dst + 2 * i) = o;
        }
        // no barrier needed here: the next block's load phase only writes W, and v is rewritten after the FFT's barriers
    }
    // the tail block (EOF): only the CTA whose range reaches the end emits it
    if (!a.rawOut && j1 == nblk) {
        int16_t *dst = a.pcm + ((int64_t)s * a.strideSamples + (int64_t)(nblk - 1) * M) * 2;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int i = tid + NT * j;
            short2 o;
            if constexpr (sizeof(T) == 8) { o.x = (short)pcm16((double)ola[j][0]); o.y = (short)pcm16((double)ola[j][1]); }
            else { o.x = (short)pcm16(ola[j][0]); o.y = (short)pcm16(ola[j][1]); }
            *reinterpret_cast<short2 *>(dst + 2 * i) = o;
        }
    }
}

}  // namespace pac
