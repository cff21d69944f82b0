// scan.cuh -- K4: the per-stream sequential part of the encoder, one warp per stream, lanes = bands/lines.
//   huffman.withdrawBits (Huffman.py:363-371) -> BitAlloc x2 with the shared reservoir (codec.py:229,257-260,
//   bitalloc.py:129-184) -> ScaleFactor per band (codec.py:273-274) -> mantissa magnitudes (quantize.py:315-342)
//   -> code length under all 10 tables, strictly-shortest wins (Huffman.py:284-308) -> depositBits
//   (codec.py:118-120) -> chunk size + running byte offset (pacfile.py:291-317).
// Block b's allocation depends on block b-1's Huffman-coded length, so this is a device-side sequential scan
// over the blocks of a stream; different streams run in different warps.
#pragma once
#include "common.cuh"

namespace pac {

struct StreamState {
    long long extraBits;     // cp.extraBits
    long long bitDeposit;    // huffman.bitDeposit
    long long outOffset;     // bytes of this stream's .pac image written so far
    long long reserved;
};

template <typename T>
struct ScanArgs {
    int S, b0, nb;                 // tile: blocks b0 .. b0+nb-1; intermediates indexed w = s*nb + (b-b0)
    int M;                         // nMDCTLines
    const int64_t *nSamples;       // [S] (NULL: every stream has exactly nb blocks, used by the per-block API)
    StreamState *state;            // [S]
    const T *lines, *smr, *bmax;
    const uint32_t *lrms;
    uint8_t *ba, *sf;              // [nwork][2][kMaxBands]
    uint8_t *tableID;              // [nwork][2]
    uint32_t *nbytes;              // [nwork][2] payload bytes
    long long *chunkOff;           // [nwork][2] offset of the chunk's length prefix in the stream image
    long long *trExtra, *trDeposit;// [nwork] optional traces of the state after each block
    const unsigned long long *lenLut;   // [kLenLutSize] 10 x 5-bit code lengths per magnitude (0 = escape)
    EncConsts ec;
    BandInfo bands;
};

// bitalloc.BitAlloc on a warp: lane b < NB owns band b.  Returns this lane's bits; *diff = bitDifference.
__device__ __forceinline__ int warp_bitalloc(double bitBudget, long long extraBits, int maxMantBits, int NB,
                                             int nLinesLane, double smrLane, uint32_t lrms, const BandInfo &bands,
                                             long long *diff) {
    const int lane = threadIdx.x & 31;
    const bool inband = lane < NB;
    int bits = 0;
    bool valid = inband;
    long long totalBits = (long long)(bitBudget + (double)extraBits);     // int() truncation, bitalloc.py:159
    while (__ballot_sync(0xffffffffu, valid) != 0u) {
        // argmax over valid bands of SMR - 6*bits, first index wins (np.argmax)
        double v = smrLane - bits * 6.;
        unsigned long long key = inband ? sortable(v) : 0ull;
        unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        unsigned hv = valid ? hi : 0u;
        unsigned mh = __reduce_max_sync(0xffffffffu, hv);
        bool c1 = valid && hi == mh;
        unsigned ml = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
        unsigned win = __ballot_sync(0xffffffffu, c1 && lo == ml);
        int iMax = __ffs(win) - 1;
        // max over ALL bands of SMR - 6*(bits-1)   (bitalloc.py:165-168)
        double v2 = smrLane - (bits - 1) * 6.;
        unsigned long long k2 = inband ? sortable(v2) : 0ull;
        unsigned h2 = (unsigned)(k2 >> 32), l2 = (unsigned)k2;
        unsigned mh2 = __reduce_max_sync(0xffffffffu, h2);
        unsigned ml2 = __reduce_max_sync(0xffffffffu, (inband && h2 == mh2) ? l2 : 0u);
        unsigned long long mk = ((unsigned long long)mh2 << 32) | ml2;
        const unsigned long long kMS = sortable(-5.0), kLR = sortable(-15.0);
        bool below = ((lrms >> iMax) & 1u) ? (mk < kMS) : (mk < kLR);
        int nl = bands.lo[iMax + 1] - bands.lo[iMax];
        bool me = lane == iMax;
        if (below && me) valid = false;
        if (totalBits - nl >= 0) {
            totalBits -= nl;
            if (me) { bits += 1; if (bits >= maxMantBits) valid = false; }
        } else if (me) valid = false;
    }
    // bits == 1 -> 0 with refund (bitalloc.py:179-180)
    unsigned ones = __ballot_sync(0xffffffffu, inband && bits == 1);
    while (ones) { int bnd = __ffs(ones) - 1; ones &= ones - 1; totalBits += bands.lo[bnd + 1] - bands.lo[bnd]; }
    if (bits == 1) bits = 0;
    *diff = totalBits - extraBits;
    (void)nLinesLane;
    return bits;
}

template <typename T, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_scan(const ScanArgs<T> a) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (s >= a.S) return;
    const int NB = a.bands.nBands, M = a.M;
    const EncConsts &ec = a.ec;
    const int largestScale = (1 << ec.nScaleBits) - 1;
    int nblkStream = a.b0 + a.nb;
    if (a.nSamples) nblkStream = (int)((a.nSamples[s] + M - 1) / M + 1);
    long long extraBits = a.state[s].extraBits, bitDeposit = a.state[s].bitDeposit, outOff = a.state[s].outOffset;
    const int nLinesLane = lane < NB ? a.bands.lo[lane + 1] - a.bands.lo[lane] : 0;
    const int bEnd = min(a.b0 + a.nb, nblkStream);

    for (int b = a.b0; b < bEnd; b++) {
        const int64_t w = (int64_t)s * a.nb + (b - a.b0);
        const uint32_t lrms = a.lrms[w];
        // withdrawBits, Huffman.py:363-371 (floor division of a positive int)
        {
            long long extra = 0;
            if (bitDeposit > 10) { extra = bitDeposit / 100; bitDeposit -= extra; }
            else if (bitDeposit < 0) { extra = bitDeposit; bitDeposit = 0; }
            extraBits += extra;                                            // codec.py:229
        }
        for (int ch = 0; ch < 2; ch++) {
            const int64_t wc = w * 2 + ch;
            double smrLane = lane < NB ? (double)a.smr[wc * kMaxBands + lane] : 0.0;
            double bmaxLane = lane < NB ? (double)a.bmax[wc * kMaxBands + lane] : 0.0;
            long long diff;
            int bits = warp_bitalloc(ec.bitBudget, extraBits, ec.maxMantBits, NB, nLinesLane, smrLane, lrms, a.bands, &diff);
            extraBits += diff;                                             // codec.py:260
            int sfl = scale_factor(bmaxLane, ec.nScaleBits, bits);       // codec.py:274
            if (lane < NB) { a.ba[wc * kMaxBands + lane] = (uint8_t)bits; a.sf[wc * kMaxBands + lane] = (uint8_t)sfl; }
            // code lengths under the 10 tables
            unsigned tot[kNTables];
#pragma unroll
            for (int t = 0; t < kNTables; t++) tot[t] = 0;
            const T *x = a.lines + wc * M;
            unsigned active = __ballot_sync(0xffffffffu, lane < NB && bits > 0);
            int nMant = 0, origin = 0;
            while (active) {
                int bd = __ffs(active) - 1;
                active &= active - 1;
                int bab = __shfl_sync(0xffffffffu, bits, bd);
                int sfb = __shfl_sync(0xffffffffu, sfl, bd);
                int lo = a.bands.lo[bd], hi = a.bands.lo[bd + 1];
                nMant += hi - lo;
                origin += bab * (hi - lo);
                for (int i = lo + lane; i < hi; i += 32) {
                    unsigned mag = mant_mag(fabs((double)x[i]), sfb, largestScale, bab);
                    unsigned long long lw = mag < (unsigned)kLenLutSize ? __ldg(a.lenLut + mag) : 0ull;
#pragma unroll
                    for (int t = 0; t < kNTables; t++) {
                        unsigned l = (unsigned)(lw >> (5 * t)) & 31u;
                        tot[t] += l ? l : (unsigned)(ec.esc_len[t] + bab);    // Huffman.py:292-298
                    }
                }
            }
            unsigned best = 0;
            int bestID = 1;
#pragma unroll
            for (int t = 0; t < kNTables; t++) {
                unsigned v = __reduce_add_sync(0xffffffffu, tot[t]);
                if (t == 0 || v < best) { best = v; bestID = t + 1; }      // Huffman.py:300-307
            }
            bitDeposit += (long long)origin - ((long long)best + nMant + ec.nTableIDBits);   // codec.py:118-120
            long long nbits = (long long)ec.fixedBits + nMant + best;       // pacfile.py:291-312
            unsigned nby = (unsigned)((nbits + 7) / 8);                     // :315-316
            if (lane == 0) {
                a.tableID[wc] = (uint8_t)bestID;
                a.nbytes[wc] = nby;
                a.chunkOff[wc] = outOff;
            }
            outOff += 4 + (long long)nby;
        }
        if (lane == 0 && a.trExtra) { a.trExtra[w] = extraBits; a.trDeposit[w] = bitDeposit; }
    }
    if (lane == 0) {
        a.state[s].extraBits = extraBits;
        a.state[s].bitDeposit = bitDeposit;
        a.state[s].outOffset = outOff;
    }
}

}  // namespace pac
