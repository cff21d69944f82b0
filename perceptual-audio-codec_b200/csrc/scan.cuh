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
    const ulonglong4 *lenLut4;          // [kLenLutSize+1] per magnitude: code length (or escape-code length) of tables 0-4 / 5-9 in
                                        // 12-bit slots, then 1 in the slot of every table that escapes this magnitude
    const uint8_t *band_of_line;        // [M]
    EncConsts ec;
    BandInfo bands;
};

// bitalloc.BitAlloc on a warp: lane b < NB owns band b.  Returns this lane's bits; *diff = bitDifference.
__device__ __forceinline__ unsigned sortable32(float v) {
    const unsigned u = __float_as_uint(v);
    return u ^ ((unsigned)((int)u >> 31) | 0x80000000u);      // negative: ~u, else u | sign bit (two instructions)
}

// fp32 fast mode: the SMRs are floats, so the arg-max keys are formed in float (one REDUX per maximum instead of two).
// The loop is the serial heart of the scan kernel (about 100 iterations per channel), so it carries its state in the form
// the next iteration needs: per lane the sortable key of SMR - 6*bits (0 once the band is invalid, so "no valid band
// left" is simply a zero maximum) and of SMR - 6*(bits-1); only the winning lane's keys change per iteration.  Line
// counts come from a shuffle of the per-lane count, the bit budget is a 32-bit integer (callers fall back to the generic
// version for budgets that do not fit).
__device__ __forceinline__ int warp_bitalloc32(int totalBits0, long long extraBits, int maxMantBits, int NB, float smrLane,
                                               uint32_t lrms, int nLinesLane, long long *diff) {
    const int lane = threadIdx.x & 31;
    const bool inband = lane < NB;
    int bits = 0;
    float fbits = 0.f;
    int totalBits = totalBits0;
    const unsigned kMS = sortable32(-5.0f), kLR = sortable32(-15.0f);
    unsigned vkey = inband ? sortable32(smrLane) : 0u;                     // valid bands only
    unsigned k2 = inband ? sortable32(smrLane + 6.f) : 0u;                 // all bands: SMR - 6*(bits-1)
    for (;;) {
        const unsigned mk1 = __reduce_max_sync(0xffffffffu, vkey);
        if (mk1 == 0u) break;                                              // no valid band left (bitalloc.py:161)
        const int iMax = __ffs(__ballot_sync(0xffffffffu, vkey == mk1)) - 1;       // first index wins (np.argmax)
        const unsigned mk2 = __reduce_max_sync(0xffffffffu, k2);
        const bool below = mk2 < (((lrms >> iMax) & 1u) ? kMS : kLR);      // :165-175
        const int nl = __shfl_sync(0xffffffffu, nLinesLane, iMax);
        const bool afford = totalBits >= nl;                               // :176
        if (afford) totalBits -= nl;
        if (lane == iMax) {
            bool valid = !below && afford;                                 // invalidated bands still get this iteration's bit (:176-178)
            if (afford) { bits += 1; fbits += 1.f; if (bits >= maxMantBits) valid = false; }
            const float v = fmaf(fbits, -6.f, smrLane);
            vkey = valid ? sortable32(v) : 0u;
            k2 = sortable32(v + 6.f);
        }
    }
    unsigned ones = __ballot_sync(0xffffffffu, inband && bits == 1);       // bits == 1 -> 0 with refund (:179-180)
    long long tb = totalBits;
    while (ones) { int bnd = __ffs(ones) - 1; ones &= ones - 1; tb += __shfl_sync(0xffffffffu, nLinesLane, bnd); }
    if (bits == 1) bits = 0;
    *diff = tb - extraBits;
    return bits;
}

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
    constexpr int LPL = 32;                      // lines per lane (M = 1024); M = 512 uses the first 16
    unsigned bandsPacked[LPL / 4];
#pragma unroll
    for (int q = 0; q < LPL / 4; q++) {
        unsigned v = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) { int i = (4 * q + r) * 32 + lane; v |= (unsigned)(i < M ? a.band_of_line[i] : 0) << (8 * r); }
        bandsPacked[q] = v;
    }
    const int lplRun = M / 32;

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
            int bits;
            const long long total0 = (long long)(ec.bitBudget + (double)extraBits);          // int() truncation, bitalloc.py:159
            if (sizeof(T) == 4 && total0 > -(1ll << 30) && total0 < (1ll << 30))
                bits = warp_bitalloc32((int)total0, extraBits, ec.maxMantBits, NB, (float)smrLane, lrms, nLinesLane, &diff);
            else bits = warp_bitalloc(ec.bitBudget, extraBits, ec.maxMantBits, NB, nLinesLane, smrLane, lrms, a.bands, &diff);
            extraBits += diff;                                             // codec.py:260
            int sfl = scale_factor(bmaxLane, ec.nScaleBits, bits);       // codec.py:274
            if (lane < NB) { a.ba[wc * kMaxBands + lane] = (uint8_t)bits; a.sf[wc * kMaxBands + lane] = (uint8_t)sfl; }
            // code lengths under the 10 tables.  Lane l owns lines 32*j + l; per line one 32-byte LUT entry gives all ten
            // lengths in 12-bit packed slots (a lane's 32 lines cannot overflow a slot), escapes add bitAlloc raw bits
            // (Huffman.py:292-298).  The 16 lines of a chunk are independent, so their loads overlap.
            const T *x = a.lines + wc * M;
            unsigned long long acc0 = 0, acc1 = 0;
#pragma unroll
            for (int j0 = 0; j0 < LPL; j0 += 16) {
                if (j0 >= lplRun) break;
                T xv[16];
#pragma unroll
                for (int j = 0; j < 16; j++) xv[j] = x[(j0 + j) * 32 + lane];
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    const int bd = (bandsPacked[(j0 + j) >> 2] >> (8 * ((j0 + j) & 3))) & 0xff;
                    const int bab = __shfl_sync(0xffffffffu, bits, bd);
                    const int sfb = __shfl_sync(0xffffffffu, sfl, bd);
                    if (bab > 0) {
                        unsigned mag = mant_mag(fabs((double)xv[j]), sfb, largestScale, bab);
                        const ulonglong2 *ep = reinterpret_cast<const ulonglong2 *>(a.lenLut4 + (mag < (unsigned)kLenLutSize ? mag : (unsigned)kLenLutSize));
                        const ulonglong2 e0 = __ldg(ep), e1 = __ldg(ep + 1);
                        acc0 += e0.x + e1.x * (unsigned long long)bab;
                        acc1 += e0.y + e1.y * (unsigned long long)bab;
                    }
                }
            }
            unsigned tot[kNTables];
#pragma unroll
            for (int t = 0; t < 5; t++) { tot[t] = (unsigned)(acc0 >> (12 * t)) & 0xfffu; tot[5 + t] = (unsigned)(acc1 >> (12 * t)) & 0xfffu; }
            int nMant = 0, origin = 0;
            {
                int nm = (lane < NB && bits > 0) ? nLinesLane : 0;
                nMant = __reduce_add_sync(0xffffffffu, nm);
                origin = __reduce_add_sync(0xffffffffu, nm * bits);
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
