// pack.cuh -- K5: one warp per (stream, block, channel) chunk re-quantises the selected lines with the
// allocation K4 fixed, Huffman-codes them with the chosen table and packs the MSB-first bit stream
// (PACFile.WriteDataBlock pacfile.py:319-351, PackedBits.WriteBits bitpack.py:36-101, StripSignBits
// codec.py:67-81, escape coding Huffman.py:292-298).  The chunk's byte offset inside the stream image was fixed by K4.
//
// Layout of a chunk (pacfile.py:324-348): [overallScale][tableID] then per band [ba-1][sf][nLines sign bits][tokens],
// then the LRMS bits.  With P(i) = number of token bits of all lines before line i (ONE running prefix over the whole
// chunk), every field's position is a per-band constant plus a value of P:
//     header of band b   S[b] + P(lo_b)            S[b] = chunk header + sum over b' < b of (band header + sign bits)
//     sign bit of line i S[b] + hdr + P(lo_b) + (i - lo_b)
//     token of line i    S[b] + hdr + signs_b + P(i)
// so the lines are walked flat, 32 per step whatever the band widths (the 17 narrowest bands hold 4..26 lines each),
// with one warp prefix sum per step; sign runs are emitted by the first lane of each band segment of a step, band
// headers by lane = band at the end.
#pragma once
#include "common.cuh"

namespace pac {

constexpr int kPackWarps = 8;
constexpr int kChunkWords = 1024;     // 4 KB per warp >= worst-case chunk (233 + 1024*30 bits = 3870 B)

template <typename T>
struct PackArgs {
    int S, b0, nb, M;
    const int64_t *nSamples;          // [S] or NULL
    const T *lines;
    const uint8_t *ba, *sf, *tableID, *oscale;
    const uint32_t *lrms;
    const uint32_t *nbytes;
    const long long *chunkOff;
    uint8_t *out;                     // [S][cap] stream images (or [nwork*2][cap] chunk buffers when perChunk)
    long long cap;
    int perChunk;                     // 1: per-block API, payload only, at out + (w*2+ch)*cap
    int *overflow;                    // [S] set to 1 when a chunk would not fit
    const uint8_t *band_of_line;      // [M]
    const uint32_t *codeLen;          // flattened tables: code << 5 | length (0: the magnitude escapes), one load per line
    int32_t *o_mant;                  // optional [nchunk][M] signed mantissa codes at line positions (pre-zeroed)
    const uint8_t *header;            // prebuilt file header (pacfile.py:237-261), numSamples patched per stream
    int headerBytes;
    EncConsts ec;
    BandInfo bands;
};

// append `len` (<= 32) bits of `val` at bit position `pos` of a zeroed big-endian word buffer
__device__ __forceinline__ void put_bits(unsigned *buf, unsigned pos, unsigned val, int len) {
    if (len <= 0) return;
    unsigned wi = pos >> 5, o = pos & 31u;
    unsigned long long v = (unsigned long long)(len == 32 ? val : (val & ((1u << len) - 1u))) << (64 - len - o);
    unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    if (hi) atomicOr(buf + wi, hi);
    if (lo) atomicOr(buf + wi + 1, lo);
}

template <typename T>
__global__ void __launch_bounds__(kPackWarps * 32, 5)
k_pack(const PackArgs<T> a) {
    __shared__ unsigned sbuf[kPackWarps][kChunkWords + 4];      // word 0: the '<L nBytes' prefix, payload from word 1
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nchunks = (int64_t)a.S * a.nb * 2;
    const int NB = a.bands.nBands, M = a.M;
    const EncConsts &ec = a.ec;
    const int largestScale = (1 << ec.nScaleBits) - 1;
    const int hdrBits = ec.nMantSizeBits + ec.nScaleBits;
    const int nGroups = M / 32;                          // M is 512 or 1024
    unsigned *buf = sbuf[warp] + 1;
    // static, lane = band: the band's first line and width
    const int loLane = lane < NB ? a.bands.lo[lane] : M;
    const int nlLane = lane < NB ? a.bands.lo[lane + 1] - a.bands.lo[lane] : 0;

    for (int64_t c = (int64_t)blockIdx.x * kPackWarps + warp; c < nchunks; c += (int64_t)gridDim.x * kPackWarps) {
        const int64_t w = c >> 1;
        const int ch = (int)(c & 1);
        const int s = nchunks <= 0xffffffffll ? (int)((uint32_t)w / (uint32_t)a.nb) : (int)(w / a.nb);
        const int b = a.b0 + (int)(w - (int64_t)s * a.nb);
        if (a.nSamples) {
            long long nblk = (a.nSamples[s] + M - 1) / M + 1;
            if (b >= nblk) continue;
        }
        const unsigned nby = a.nbytes[c];
        const unsigned nwords = (nby + 3) >> 2;
        if (nwords > (unsigned)kChunkWords) { if (lane == 0 && a.overflow) a.overflow[s] = 1; continue; }
        for (unsigned i = lane; i < nwords + 2; i += 32) buf[i] = 0;
        const int tid = a.tableID[c] - 1;
        const int off_t = ec.off[tid], nkeys_t = ec.nkeys[tid];
        const unsigned escc = ec.esc_code[tid];
        const int escl = ec.esc_len[tid];
        const T *x = a.lines + c * M;
        // lane = band: allocation, scale factor, and S[b] (exclusive prefix of band header + sign bits)
        const int babL = lane < NB ? a.ba[c * kMaxBands + lane] : 0, sfL = lane < NB ? a.sf[c * kMaxBands + lane] : 0;
        const unsigned pbL = (unsigned)babL | (unsigned)sfL << 8;
        const int fixedL = lane < NB ? hdrBits + (babL ? nlLane : 0) : 0;
        int inclS = fixedL;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inclS, o); if (lane >= o) inclS += v; }
        const int pos0 = ec.nScaleBits + ec.nTableIDBits;
        const int SL = pos0 + inclS - fixedL;                              // S[b]
        const int CsgnL = SL + hdrBits, CtokL = SL + fixedL;               // sign run / token area of band b, before adding P
        const int fixedTotal = pos0 + __shfl_sync(0xffffffffu, inclS, 31);
        int PbandL = -1;                                                   // P(lo_b), filled in when the walk reaches the band
        __syncwarp();
        int carry = 0;                                                     // token bits of all lines before this step
        // the next step's line and band index are requested one step ahead (two registers): a step's own work is too short to cover the
        // round trip (ncu: 22 % of the kernel's stall samples sat on the first use of x[i])
        T xn = x[lane];
        int bn = a.band_of_line[lane];
#pragma unroll 1
        for (int g = 0; g < nGroups; g++) {
            const int i0 = 32 * g, i = i0 + lane;
            const int bd = bn;
            const T xv = xn;
            if (g + 1 < nGroups) { xn = x[i + 32]; bn = a.band_of_line[i + 32]; }
            const unsigned pb = __shfl_sync(0xffffffffu, pbL, bd);
            const int bab = (int)(pb & 0xff), sfb = (int)(pb >> 8);
            unsigned code = 0, code2 = 0;
            int len = 0, len2 = 0;
            bool sg = false;
            if (bab) {
                sg = signbit((double)xv);                                  // np.signbit semantics (quantize.py:333): -0.0 counts
                const unsigned mag = mant_mag(fabs((double)xv), sfb, largestScale, bab);
                if (a.o_mant) a.o_mant[c * M + i] = (int32_t)(mag + (sg ? (1u << (bab - 1)) : 0u));
                const unsigned cl = (int)mag < nkeys_t ? __ldg(a.codeLen + off_t + mag) : 0u;
                const int l = (int)(cl & 31u);
                if (l) { code = cl >> 5; len = l; }
                else { code = escc; len = escl; code2 = mag; len2 = bab; }      // Huffman.py:296-298
            }
            int tl = len + len2;
            // an escape token (escape code + raw magnitude) is written as ONE field when it fits 32 bits (always with the
            // stock tables: escape codes are <= 13 bits, magnitudes <= 16)
            if (len2 && tl <= 32) { code = (code << len2) | code2; len = tl; len2 = 0; }
            int incl = tl;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const int P = carry + incl - tl;                                // P(i)
            // bands that start inside this step learn their P(lo_b)
            {
                const int v = __shfl_sync(0xffffffffu, P, (loLane - i0) & 31);
                if (loLane >= i0 && loLane < i0 + 32) PbandL = v;
            }
            // tokens (:337-341)
            const unsigned tp = (unsigned)(__shfl_sync(0xffffffffu, CtokL, bd) + P);
            put_bits(buf, tp, code, len);
            if (len2) put_bits(buf, tp + len, code2, len2);
            // sign bits of the whole band come first (:335-336): the first lane of each band segment of this step writes the run
            const unsigned bal = __ballot_sync(0xffffffffu, sg);
            const int bdPrev = __shfl_up_sync(0xffffffffu, bd, 1);
            const int loB = __shfl_sync(0xffffffffu, loLane, bd), nlB = __shfl_sync(0xffffffffu, nlLane, bd);
            const int csg = __shfl_sync(0xffffffffu, CsgnL, bd), pband = __shfl_sync(0xffffffffu, PbandL, bd);
            if (bab && (lane == 0 || bd != bdPrev)) {
                int n = loB + nlB - i;                                      // lines of this band from i on ...
                if (n > 32 - lane) n = 32 - lane;                           // ... that belong to this step
                const unsigned run = __brev(bal >> lane) >> (32 - n);       // line i first (MSB-first stream)
                put_bits(buf, (unsigned)(csg + pband + (i - loB)), run, n);
            }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane < NB && PbandL < 0) PbandL = carry;                        // bands without lines
        // band headers (:329-332), lane = band
        if (lane < NB) {
            const unsigned baF = (unsigned)(babL ? babL - 1 : 0) & ((1u << ec.nMantSizeBits) - 1u);
            const unsigned sfF = (unsigned)sfL & ((1u << ec.nScaleBits) - 1u);
            put_bits(buf, (unsigned)(SL + PbandL), (baF << ec.nScaleBits) | sfF, hdrBits);
        }
        if (lane == 0) {
            put_bits(buf, 0, a.oscale[c], ec.nScaleBits);                       // pacfile.py:324
            put_bits(buf, ec.nScaleBits, (unsigned)(tid + 1), ec.nTableIDBits);  // :326
            // LRMS: band 0 first (:347-348) -> bit-reverse the mask into MSB-first order
            put_bits(buf, (unsigned)(fixedTotal + carry), __brev(a.lrms[w]) >> (32 - NB), NB);
        }
        __syncwarp();
        // ---- store: [<L nBytes][payload] at the offset K4 assigned, as ONE byte stream out of shared memory (word -1 holds the
        // length prefix, byte-swapped so that the big-endian stream reads as '<L'): unaligned head and tail bytewise, the body as
        // 16-byte stores (512 contiguous bytes per warp instruction -- the output may be pinned HOST memory written across PCIe,
        // where 32-byte transactions would waste the link)
        uint8_t *dst;
        const unsigned *src;
        unsigned nout;
        if (a.perChunk) {
            dst = a.out + c * a.cap; src = buf; nout = nby;
            if (a.cap < (long long)nby) { if (lane == 0 && a.overflow) a.overflow[s] = 1; continue; }
        } else {
            const long long off = a.chunkOff[c];
            if (a.cap - off < 4 + (long long)nby) { if (lane == 0 && a.overflow) a.overflow[s] = 1; continue; }
            dst = a.out + (long long)s * a.cap + off; src = buf - 1; nout = nby + 4;
            if (lane == 0) buf[-1] = __byte_perm(nby, 0, 0x0123);                  // '<L', pacfile.py:317
            if (b == 0 && ch == 0) {                                              // WriteFileHeader, pacfile.py:237-261
                uint8_t *h = a.out + (long long)s * a.cap;
                for (int i = lane; i < a.headerBytes; i += 32) h[i] = a.header[(long long)s * a.headerBytes + i];
            }
            __syncwarp();
        }
        auto byteAt = [&](unsigned i) { return (uint8_t)(src[i >> 2] >> (24 - 8 * (i & 3))); };
        const unsigned head = min(nout, (unsigned)((16 - ((uintptr_t)dst & 15)) & 15));
        if (lane < head) dst[lane] = byteAt(lane);
        const unsigned nvec = (nout - head) >> 4;
        const unsigned sh = 8 * (head & 3);
        for (unsigned v = lane; v < nvec; v += 32) {
            const unsigned wi = (head >> 2) + 4 * v;                               // first source word of this 16-byte group
            const unsigned w0 = src[wi], w1 = src[wi + 1], w2 = src[wi + 2], w3 = src[wi + 3], w4 = src[wi + 4];
            uint4 o;
            o.x = __byte_perm(__funnelshift_l(w1, w0, sh), 0, 0x0123);
            o.y = __byte_perm(__funnelshift_l(w2, w1, sh), 0, 0x0123);
            o.z = __byte_perm(__funnelshift_l(w3, w2, sh), 0, 0x0123);
            o.w = __byte_perm(__funnelshift_l(w4, w3, sh), 0, 0x0123);
            *reinterpret_cast<uint4 *>(dst + head + 16 * v) = o;
        }
        const unsigned done = head + 16 * nvec;
        if (done + lane < nout) dst[done + lane] = byteAt(done + lane);
        __syncwarp();
    }
}

// ---------------------------------------------------------------- K5b: drain a tile's bytes to the caller's pinned host image
// Host-buffer encode with a PINNED `out`: k_pack could write every chunk straight across PCIe (and does when the batch is staged in
// several groups), but its bursts (6 GB of a step's output inside the ~13 % of the step it runs) then queue behind the link and hold
// the SM resources of a 2000-warp kernel while they wait: pack 200 -> 251 ms, analysis 1244 -> 1283 ms per 4096 x 60 s.  Instead k_pack
// writes into a device image and this kernel -- a few warps, one stream per warp at a time -- copies the byte range every stream gained
// in the tile ([tileBeg, tileEnd), recorded by k_scan) to the same offsets of the host image, at the link's pace, beside the next
// tile's analysis.  Device image and host image have the same alignment modulo 16, so the body moves as 16-byte words.
struct DrainArgs {
    const uint8_t *img;              // [S][cap] device image
    uint8_t *dst;                    // [S][cap] device-visible alias of the pinned host image
    long long cap;
    const long long *tileBeg, *tileEnd;   // [S]
    int S;
};
constexpr int kDrainWarps = 4;
__global__ void __launch_bounds__(kDrainWarps * 32) k_drain(const DrainArgs a) {
    const int lane = threadIdx.x & 31;
    const int nw = gridDim.x * kDrainWarps;
    for (int s = blockIdx.x * kDrainWarps + (threadIdx.x >> 5); s < a.S; s += nw) {
        long long b = a.tileBeg[s], e = a.tileEnd[s];
        if (e > a.cap) e = a.cap;                                   // an overflowing stream is reported through outBytes; nothing past its row
        if (e <= b) continue;
        const uint8_t *src = a.img + (long long)s * a.cap;
        uint8_t *d = a.dst + (long long)s * a.cap;
        long long b16 = b + (long long)((16 - ((uintptr_t)(d + b) & 15)) & 15);
        if (b16 > e) b16 = e;
        for (long long i = b + lane; i < b16; i += 32) d[i] = src[i];
        const long long e16 = b16 + ((e - b16) & ~15ll);
        long long i = b16 + 16 * lane;
        for (; i + 3 * 512 < e16; i += 4 * 512) {                   // four 16-byte words in flight per lane
            const uint4 v0 = *reinterpret_cast<const uint4 *>(src + i), v1 = *reinterpret_cast<const uint4 *>(src + i + 512);
            const uint4 v2 = *reinterpret_cast<const uint4 *>(src + i + 1024), v3 = *reinterpret_cast<const uint4 *>(src + i + 1536);
            *reinterpret_cast<uint4 *>(d + i) = v0; *reinterpret_cast<uint4 *>(d + i + 512) = v1;
            *reinterpret_cast<uint4 *>(d + i + 1024) = v2; *reinterpret_cast<uint4 *>(d + i + 1536) = v3;
        }
        for (; i < e16; i += 512) *reinterpret_cast<uint4 *>(d + i) = *reinterpret_cast<const uint4 *>(src + i);
        for (long long j = e16 + lane; j < e; j += 32) d[j] = src[j];
    }
}

}  // namespace pac
