// pack.cuh -- K5: one warp per (stream, block, channel) chunk re-quantises the selected lines with the
// allocation K4 fixed, Huffman-codes them with the chosen table and packs the MSB-first bit stream
// (PACFile.WriteDataBlock pacfile.py:319-351, PackedBits.WriteBits bitpack.py:36-101, StripSignBits
// codec.py:67-81, escape coding Huffman.py:292-298).  Bit positions inside a chunk come from a warp-level
// exclusive prefix sum of the code lengths; the chunk's byte offset inside the stream image was fixed by K4.
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
    const uint32_t *codeLut;          // flattened code values   (PacHuffTables.code)
    const uint8_t *lenLutFlat;        // flattened code lengths  (PacHuffTables.len)
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
__global__ void __launch_bounds__(kPackWarps * 32)
k_pack(const PackArgs<T> a) {
    __shared__ unsigned sbuf[kPackWarps][kChunkWords + 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t nchunks = (int64_t)a.S * a.nb * 2;
    const int NB = a.bands.nBands, M = a.M;
    const EncConsts &ec = a.ec;
    const int largestScale = (1 << ec.nScaleBits) - 1;
    unsigned *buf = sbuf[warp];

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
        __syncwarp();
        const int tid = a.tableID[c] - 1;
        const int off_t = ec.off[tid], nkeys_t = ec.nkeys[tid];
        const unsigned escc = ec.esc_code[tid];
        const int escl = ec.esc_len[tid];
        const T *x = a.lines + c * M;
        unsigned pos = 0;
        if (lane == 0) {
            put_bits(buf, 0, a.oscale[c], ec.nScaleBits);                       // pacfile.py:324
            put_bits(buf, ec.nScaleBits, (unsigned)(tid + 1), ec.nTableIDBits);  // :326
        }
        pos = ec.nScaleBits + ec.nTableIDBits;
        for (int bd = 0; bd < NB; bd++) {
            const int bab = a.ba[c * kMaxBands + bd], sfb = a.sf[c * kMaxBands + bd];
            if (lane == 0) {
                put_bits(buf, pos, (unsigned)(bab ? bab - 1 : 0), ec.nMantSizeBits);   // :329-331
                put_bits(buf, pos + ec.nMantSizeBits, (unsigned)sfb, ec.nScaleBits);   // :332
            }
            pos += ec.nMantSizeBits + ec.nScaleBits;
            if (!bab) continue;
            const int lo = a.bands.lo[bd], hi = a.bands.lo[bd + 1];
            // sign bits of the whole band first (:335-336), np.signbit semantics (quantize.py:333)
            for (int i0 = lo; i0 < hi; i0 += 32) {
                int i = i0 + lane;
                bool sg = false;
                if (i < hi) sg = signbit((double)x[i]);
                unsigned bal = __ballot_sync(0xffffffffu, sg);
                int n = min(32, hi - i0);
                if (lane == 0) put_bits(buf, pos, __brev(bal) >> (32 - n), n);
                pos += n;
            }
            // then the Huffman tokens (:337-341)
            for (int i0 = lo; i0 < hi; i0 += 32) {
                int i = i0 + lane;
                unsigned code = 0;
                int len = 0, len2 = 0;
                unsigned code2 = 0;
                if (i < hi) {
                    unsigned mag = mant_mag(fabs((double)x[i]), sfb, largestScale, bab);
                    if (a.o_mant) a.o_mant[c * M + i] = (int32_t)(mag + (signbit((double)x[i]) ? (1u << (bab - 1)) : 0u));
                    int l = (int)mag < nkeys_t ? a.lenLutFlat[off_t + mag] : 0;
                    if (l) { code = a.codeLut[off_t + mag]; len = l; }
                    else { code = escc; len = escl; code2 = mag; len2 = bab; }      // Huffman.py:296-298
                }
                int tl = len + len2;
                // an escape token (escape code + raw magnitude, Huffman.py:296-298) is written as ONE field when it fits 32 bits
                // (always with the stock tables: escape codes are <= 13 bits, magnitudes <= 16)
                if (len2 && tl <= 32) { code = (code << len2) | code2; len = tl; len2 = 0; }
                int incl = tl;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                unsigned p = pos + (unsigned)(incl - tl);
                put_bits(buf, p, code, len);
                if (len2) put_bits(buf, p + len, code2, len2);
                pos += (unsigned)__shfl_sync(0xffffffffu, incl, 31);
            }
        }
        // LRMS: band 0 first (:347-348) -> bit-reverse the mask into MSB-first order
        if (lane == 0) put_bits(buf, pos, __brev(a.lrms[w]) >> (32 - NB), NB);
        __syncwarp();
        // ---- store: [<L nBytes][payload] at the offset K4 assigned
        uint8_t *dst;
        long long room;
        if (a.perChunk) { dst = a.out + c * a.cap; room = a.cap; }
        else {
            long long off = a.chunkOff[c];
            dst = a.out + (long long)s * a.cap + off;
            room = a.cap - off;
            if (room < 4 + (long long)nby) { if (lane == 0 && a.overflow) a.overflow[s] = 1; continue; }
            if (lane < 4) dst[lane] = (uint8_t)(nby >> (8 * lane));               // '<L', pacfile.py:317
            dst += 4;
            if (b == 0 && ch == 0) {                                              // WriteFileHeader, pacfile.py:237-261
                uint8_t *h = a.out + (long long)s * a.cap;
                for (int i = lane; i < a.headerBytes; i += 32) h[i] = a.header[(long long)s * a.headerBytes + i];
            }
        }
        if (a.perChunk && room < (long long)nby) { if (lane == 0 && a.overflow) a.overflow[s] = 1; continue; }
        for (unsigned i = lane; i < nby; i += 32) dst[i] = (uint8_t)(buf[i >> 2] >> (24 - 8 * (i & 3)));
        __syncwarp();
    }
}

}  // namespace pac
