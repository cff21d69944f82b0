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
    long long *tileBeg, *tileEnd;  // [S] optional: byte range [beg, end) of the stream's image this tile produces (k_drain); the tile
                                   //     with block 0 starts at 0 (k_pack writes the file header with block 0)
    const unsigned long long *lenLut;   // [kLenLutSize] 10 x 5-bit code lengths per magnitude (0 = escape)
    const ulonglong4 *lenLut4;          // [kLenLutSize+1] per magnitude: code length (or escape-code length) of tables 0-4 / 5-9 in
                                        // 12-bit slots, then 1 in the slot of every table that escapes this magnitude
    const uint8_t *band_of_line;        // [M]
    EncConsts ec;
    BandInfo bands;
};

// ------------------------------------------------------------------------------------------------------------------
// bitalloc.BitAlloc (bitalloc.py:129-184) on a warp: lane b < NB owns band b.
//
// The reference hands out one bit per iteration (~100-150 iterations per channel, each a pair of 25-way maxima): that
// loop WAS the serial heart of the reservoir chain.  Its greedy arg-max over SMR_b - 6*bits_b is a merge of 25 strictly
// decreasing sequences, so as long as nothing happens the state after every entry with key >= tau has been served is
// known in closed form (bits_b = number of band b's keys >= tau).  warp_bitalloc_jump advances by such cuts and only
// walks single iterations through the +-0.5 dB zones around the stop thresholds, where the outcome depends on rounding:
//   * budget: tau is raised by bisection on the EXACT cost until the whole cut is affordable;
//   * max-NMR stop rule (:163-168): per jump a band class (M/S, L/R) is NORMAL (the rule provably does not fire) or
//     TERMINAL (it provably fires: the band takes one more bit -- allocate-after-invalidate -- and leaves);
//   * a band wider than the remaining budget can only ever be invalidated (the budget never grows in the loop): dropped.
// tests/model_bitalloc.py is the numpy statement of exactly this control flow; tests/test_model.py proves it equal to the
// plain loop on random and adversarial problems (ties on the 6 dB lattice, thresholds hit exactly, reservoirs from
// negative to 500 k bits) in both arithmetics.  Typical: 2-3 jumps + 3-4 single iterations per channel.
//
// T = double: the reference's arithmetic (fp64 verification mode, bit-exact).  T = float: the fast mode's keys
// (key = fma(bits, -6, SMR), NMR = key + 6).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned sortable32(float v) {
    const unsigned u = __float_as_uint(v);
    return u ^ ((unsigned)((int)u >> 31) | 0x80000000u);      // negative: ~u, else u | sign bit (two instructions)
}
__device__ __forceinline__ float unsortable32(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
// warp maximum; lanes with on == false do not take part; -inf when nobody does
__device__ __forceinline__ float wmax_on(float v, bool on) {
    const unsigned k = __reduce_max_sync(0xffffffffu, on ? sortable32(v) : 0u);
    return k ? unsortable32(k) : -INFINITY;
}
__device__ __forceinline__ double wmax_on(double v, bool on) {
    const unsigned long long key = on ? sortable(v) : 0ull;
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    const unsigned long long mk = ((unsigned long long)mh << 32) | ml;
    if (mk == 0ull) return -INFINITY;
    return __longlong_as_double((long long)((mk >> 63) ? (mk & 0x7fffffffffffffffull) : ~mk));
}
__device__ __forceinline__ float ba_key(float smr, int bits) { return fmaf((float)bits, -6.f, smr); }
__device__ __forceinline__ double ba_key(double smr, int bits) { return smr - bits * 6.; }
__device__ __forceinline__ float ba_nmr(float smr, int bits) { return ba_key(smr, bits) + 6.f; }          // fast mode: key + 6
__device__ __forceinline__ double ba_nmr(double smr, int bits) { return smr - (bits - 1) * 6.; }         // bitalloc.py:165

// largest n in [lo, cap] with key(j) >= tau for all lo <= j < n: closed-form guess, fixed up with the loop's own key expression
template <typename T>
__device__ __forceinline__ int ba_count_ge(T smr, T tau, int lo, int cap) {
    T g = floor((smr - tau) * (T)(1.0 / 6.0));
    g = g < (T)-1 ? (T)-1 : (g > (T)64 ? (T)64 : g);          // also maps tau = -inf (g = +inf) to "everything"
    int n = (int)g + 1;
    n = n < lo ? lo : n;
    n = n > cap ? cap : n;
    // the guess is off by at most one (one rounding in the division): a single corrective step (tests/model_bitalloc.py asserts it)
    const bool dec = n > lo && ba_key(smr, n - 1) < tau;
    const bool inc = !dec && n < cap && ba_key(smr, n) >= tau;
    return n - (dec ? 1 : 0) + (inc ? 1 : 0);
}

template <typename T>
__device__ __forceinline__ int warp_bitalloc_jump(long long total0, long long extraBits, int maxMantBits, int NB, T smrLane,
                                                  uint32_t lrms, int nLinesLane, long long *diff) {
    const int lane = threadIdx.x & 31;
    const bool inband = lane < NB;
    const bool isMS = (lrms >> lane) & 1u;
    const T NINF = (T)-INFINITY;
    int bits = 0;
    bool valid = inband;
    // int(bitBudget + extraBits), :159.  A whole channel costs at most 16 x 1024 bits, so any budget beyond +-2^30 behaves like 2^30:
    // the loop runs on a clamped 32-bit budget and the difference is put back at the end
    const long long clampLo = -(1ll << 30), clampHi = 1ll << 30;
    const int total00 = (int)(total0 < clampLo ? clampLo : (total0 > clampHi ? clampHi : total0));
    int total = total00;
    for (;;) {
        valid = valid && nLinesLane <= total;                              // prune
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (!vmask) break;                                                 // :161
        const T kcur = ba_key(smrLane, bits);
        const T m = wmax_on(kcur, valid);                                  // the next pick's key
        const T F = wmax_on(ba_nmr(smrLane, bits), inband && !valid);      // max NMR held by bands that have left
        // ---- classify the two band classes for an event-free advance
        const unsigned msMask = vmask & lrms, lrMask = vmask & ~lrms;
        T flo = NINF;
        bool termMS = false, termLR = false, ok = true;
        if (msMask) {
            if (m >= (T)-10.5) flo = (T)-10.5;
            else if (m < (T)-11.5 && F < (T)-5.5) termMS = true;
            else ok = false;
        }
        if (ok && lrMask) {
            const T k1 = termMS ? wmax_on(kcur, valid && isMS) : NINF;     // first M/S band to leave in this jump
            const T FB = F > k1 ? F : k1;
            if (FB >= (T)-14.5) {}                                         // a band that has left keeps max NMR >= -15: L/R rule silent
            else if (m >= (T)-20.5) flo = flo > (T)-20.5 ? flo : (T)-20.5;
            else if (m < (T)-21.5 && FB < (T)-15.5) termLR = true;
            else ok = false;
        }
        if (ok && flo < m) {
            const bool term = isMS ? termMS : termLR;
            int cap = valid ? (term ? bits + 1 : maxMantBits) : bits;
            cap = cap > maxMantBits ? maxMantBits : cap;
            int nb = ba_count_ge<T>(smrLane, flo, bits, cap);
            int cost = __reduce_add_sync(0xffffffffu, (nb - bits) * nLinesLane);
            if (cost > total) {
                T hi = m + (T)1;                                           // nothing is >= hi
                T lo = flo;
                if (flo == NINF) lo = -wmax_on(-ba_key(smrLane, maxMantBits - 1), valid) - (T)1;
                nb = bits; cost = 0;
                // six probes: each further one would cost as much as the single iteration it saves (tests/model_bitalloc.py)
#pragma unroll 1
                for (int it = 0; it < 6; it++) {
                    const T mid = (lo + hi) * (T)0.5;
                    const int n2 = ba_count_ge<T>(smrLane, mid, bits, cap);
                    const int c2 = __reduce_add_sync(0xffffffffu, (n2 - bits) * nLinesLane);
                    if (c2 <= total) { hi = mid; nb = n2; cost = c2; }
                    else lo = mid;
                }
            }
            if (__ballot_sync(0xffffffffu, nb != bits)) {
                total -= cost;
                if (nb >= maxMantBits || (term && nb > bits)) valid = false;
                bits = nb;
                continue;
            }
        }
        // ---- one exact iteration (:162-176)
        const int iMax = __ffs(__ballot_sync(0xffffffffu, valid && kcur == m)) - 1;   // first index wins (np.argmax)
        const T mv = wmax_on(ba_nmr(smrLane, bits), valid);
        const T mxAll = F > mv ? F : mv;                                               // max over ALL bands (:165)
        const bool below = mxAll < (((lrms >> iMax) & 1u) ? (T)-5 : (T)-15);
        const int nl = __shfl_sync(0xffffffffu, nLinesLane, iMax);
        const bool afford = total >= nl;                                               // :172
        if (afford) total -= nl;
        if (lane == iMax) {
            if (below) valid = false;                                                  // still takes this iteration's bit
            if (afford) { bits += 1; if (bits >= maxMantBits) valid = false; }
            else valid = false;
        }
    }
    // bits == 1 -> 0 with refund (:179-180)
    total += __reduce_add_sync(0xffffffffu, (inband && bits == 1) ? nLinesLane : 0);
    if (bits == 1) bits = 0;
    *diff = (total0 - extraBits) + (long long)(total - total00);          // totalBits - extraBits (:182), the clamp put back
    return bits;
}

// One CTA per stream, WARPS warps (even).  Warp 0 carries the stream's serial state (reservoir, savings pool, byte
// offset) and runs the two BitAllocs of a block; then ALL warps share the line phase (half of the warps per channel, each
// a contiguous range of lines: quantise, look up the ten code lengths), whose loads were issued before the BitAllocs so
// that their latency hides behind them; warp 0 finally picks the table and updates the state.  Two CTA barriers per block.
// WARPS == 1 (many streams in flight: the kernel is a throughput problem and what a stream costs is its register-time): ONE warp
// per stream does everything, both channels' line phases one after the other -- no partner warps idling at the barrier through the
// BitAllocs (60 % of a block's latency); kScanSoloStreams such warps share a CTA and never synchronise with each other.
constexpr int kScanSoloStreams = 4;
#ifndef PAC_SCAN_SOLO_MINB
#define PAC_SCAN_SOLO_MINB 8         // resident CTAs per SM the one-warp variant is compiled for (8 x 128 threads: 64 registers)
#endif
template <typename T, int WARPS>
__global__ void __launch_bounds__(WARPS == 1 ? kScanSoloStreams * 32 : WARPS * 32, WARPS == 1 ? PAC_SCAN_SOLO_MINB : 1)
k_scan(const ScanArgs<T> a) {
    static_assert(WARPS == 1 || (WARPS >= 2 && WARPS % 2 == 0), "half of the warps per channel");
    constexpr bool SOLO = WARPS == 1;
    constexpr int WPC = SOLO ? 1 : WARPS / 2;    // warps per channel
    constexpr int LPLMAX = 1024 / (32 * WPC);    // lines per lane (and channel) at M = 1024
    constexpr int SPC = SOLO ? kScanSoloStreams : 1;      // streams per CTA
    __shared__ unsigned short sBitsSfAll[SPC][2][kMaxBands];       // bits | sf << 8 per (channel, band) of the current block
    __shared__ unsigned sTotAll[SPC][SOLO ? 2 : WARPS][kNTables];  // per-warp (SOLO: per-channel) totals of the ten table lengths
    const int lane = threadIdx.x & 31, warp = SOLO ? 0 : (int)(threadIdx.x >> 5);
    const int slot = SOLO ? (int)(threadIdx.x >> 5) : 0;
    auto &sBitsSf = sBitsSfAll[slot];
    auto &sTot = sTotAll[slot];
    auto sync = [&]() { if constexpr (SOLO) __syncwarp(); else __syncthreads(); };
    const int s = SOLO ? (int)blockIdx.x * SPC + slot : (int)blockIdx.x;
    if (s >= a.S) return;
    const int NB = a.bands.nBands, M = a.M;
    const EncConsts &ec = a.ec;
    const int largestScale = (1 << ec.nScaleBits) - 1;
    int nblkStream = a.b0 + a.nb;
    if (a.nSamples) nblkStream = (int)((a.nSamples[s] + M - 1) / M + 1);
    long long extraBits = 0, bitDeposit = 0, outOff = 0;
    if (warp == 0) {
        extraBits = a.state[s].extraBits; bitDeposit = a.state[s].bitDeposit; outOff = a.state[s].outOffset;
        if (lane == 0 && a.tileBeg) a.tileBeg[s] = a.b0 == 0 ? 0ll : outOff;
    }
    const int nLinesLane = lane < NB ? a.bands.lo[lane + 1] - a.bands.lo[lane] : 0;
    const int bEnd = min(a.b0 + a.nb, nblkStream);
    // line phase geometry: this warp owns lines [l0, l0 + LPW) of channel myCh; lane l owns l0 + 32*j + l
    const int myCh = warp / WPC, LPW = M / WPC, l0 = (warp % WPC) * LPW, lpl = LPW / 32;
    unsigned bandsPacked[(LPLMAX + 3) / 4];
#pragma unroll
    for (int q = 0; q < (LPLMAX + 3) / 4; q++) {
        unsigned v = 0;
#pragma unroll
        for (int r = 0; r < 4; r++) { const int j = 4 * q + r; v |= (unsigned)(j < lpl ? a.band_of_line[l0 + j * 32 + lane] : 0) << (8 * r); }
        bandsPacked[q] = v;
    }

    // warp 0: band inputs of a block are fetched one block ahead (their latency would otherwise sit on the serial chain)
    uint32_t lrmsN = 0;
    T smrN[2] = {0, 0}, bmaxN[2] = {0, 0};
    auto fetchBands = [&](int b) {
        const int64_t w = (int64_t)s * a.nb + (b - a.b0);
        lrmsN = a.lrms[w];
#pragma unroll
        for (int ch = 0; ch < 2; ch++) {
            smrN[ch] = lane < NB ? a.smr[(w * 2 + ch) * kMaxBands + lane] : (T)0;
            bmaxN[ch] = lane < NB ? a.bmax[(w * 2 + ch) * kMaxBands + lane] : (T)0;
        }
    };
    if (warp == 0 && a.b0 < bEnd) fetchBands(a.b0);

    for (int b = a.b0; b < bEnd; b++) {
        const int64_t w = (int64_t)s * a.nb + (b - a.b0);
        // every warp: the first chunk of its lines of this block (they do not depend on the allocation), so that their
        // latency hides behind the BitAllocs; further chunks (fewer warps per stream) are fetched one chunk ahead below
        constexpr int CH = WARPS >= 8 ? (LPLMAX < 8 ? LPLMAX : 8) : 4;   // lines per lane per chunk (fewer in flight where registers count)
        const T *xrow0 = a.lines + (w * 2 + myCh) * M + l0 + lane;
        T xv[CH];
#pragma unroll
        for (int j = 0; j < CH; j++) xv[j] = j < lpl ? xrow0[j * 32] : (T)0;
        if (warp == 0) {
            const uint32_t lrms = lrmsN;
            const T smrL[2] = {smrN[0], smrN[1]}, bmaxL[2] = {bmaxN[0], bmaxN[1]};
            if (b + 1 < bEnd) fetchBands(b + 1);
            // withdrawBits, Huffman.py:363-371 (floor division of a positive int)
            {
                long long extra = 0;
                if (bitDeposit > 10) { extra = bitDeposit / 100; bitDeposit -= extra; }
                else if (bitDeposit < 0) { extra = bitDeposit; bitDeposit = 0; }
                extraBits += extra;                                            // codec.py:229
            }
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                const int64_t wc = w * 2 + ch;
                long long diff;
                const long long total0 = (long long)(ec.bitBudget + (double)extraBits);          // int() truncation, bitalloc.py:159
                const int bits = warp_bitalloc_jump<T>(total0, extraBits, ec.maxMantBits, NB, smrL[ch], lrms, nLinesLane, &diff);
                extraBits += diff;                                             // codec.py:260
                const int sfl = scale_factor((double)bmaxL[ch], ec.nScaleBits, bits);           // codec.py:274
                if (lane < NB) {
                    a.ba[wc * kMaxBands + lane] = (uint8_t)bits; a.sf[wc * kMaxBands + lane] = (uint8_t)sfl;
                    sBitsSf[ch][lane] = (unsigned short)(bits | (sfl << 8));
                }
            }
        }
        sync();
        // ---- line phase: code lengths under the 10 tables.  Per line one 32-byte LUT entry gives all ten lengths in 12-bit
        // packed slots (a lane's <= 32 lines cannot overflow a slot), escapes add bitAlloc raw bits (Huffman.py:292-298).
#pragma unroll 1
        for (int pass = 0; pass < (SOLO ? 2 : 1); pass++) {
            const int chL = SOLO ? pass : myCh;                  // the channel whose lines this pass covers
            const T *xrow = xrow0 + (SOLO ? pass * M : 0);
            unsigned long long acc0 = 0, acc1 = 0;
#pragma unroll
            for (int c0 = 0; c0 < LPLMAX; c0 += CH) {
                if (c0 >= lpl) break;
                T xn[CH];
                if constexpr (SOLO) {                            // next chunk of this channel, or (after channel 0's last) channel 1's first
                    const bool more = c0 + CH < lpl;
                    const T *nx = more ? xrow + (c0 + CH) * 32 : xrow + M;
#pragma unroll
                    for (int j = 0; j < CH; j++) xn[j] = (more ? c0 + CH + j < lpl : (pass == 0 && j < lpl)) ? nx[j * 32] : (T)0;
                } else if (c0 + CH < LPLMAX) {
#pragma unroll
                    for (int j = 0; j < CH; j++) xn[j] = c0 + CH + j < lpl ? xrow[(c0 + CH + j) * 32] : (T)0;
                }
#pragma unroll
                for (int jj = 0; jj < CH; jj++) {
                    const int j = c0 + jj;
                    if (j < lpl) {
                        const int bd = (bandsPacked[j >> 2] >> (8 * (j & 3))) & 0xff;
                        const unsigned bs = sBitsSf[chL][bd];
                        const int bab = (int)(bs & 0xffu), sfb = (int)(bs >> 8);
                        if (bab > 0) {
                            unsigned mag = mant_mag(fabs((double)xv[jj]), sfb, largestScale, bab);
                            const ulonglong2 *ep = reinterpret_cast<const ulonglong2 *>(a.lenLut4 + (mag < (unsigned)kLenLutSize ? mag : (unsigned)kLenLutSize));
                            const ulonglong2 e0 = __ldg(ep), e1 = __ldg(ep + 1);
                            acc0 += e0.x + e1.x * (unsigned long long)bab;
                            acc1 += e0.y + e1.y * (unsigned long long)bab;
                        }
                    }
                }
                if (c0 + CH < LPLMAX || SOLO) {
#pragma unroll
                    for (int j = 0; j < CH; j++) xv[j] = xn[j];
                }
            }
#pragma unroll
            for (int t = 0; t < 5; t++) {
                const unsigned v0 = __reduce_add_sync(0xffffffffu, (unsigned)(acc0 >> (12 * t)) & 0xfffu);
                const unsigned v1 = __reduce_add_sync(0xffffffffu, (unsigned)(acc1 >> (12 * t)) & 0xfffu);
                if (lane == 0) { sTot[SOLO ? pass : warp][t] = v0; sTot[SOLO ? pass : warp][5 + t] = v1; }
            }
        }
        sync();
        if (warp == 0) {
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
                const int64_t wc = w * 2 + ch;
                const unsigned bs = lane < NB ? sBitsSf[ch][lane] : 0u;
                const int bits = (int)(bs & 0xffu);
                const int nm = bits > 0 ? nLinesLane : 0;
                const int nMant = __reduce_add_sync(0xffffffffu, nm);
                const int origin = __reduce_add_sync(0xffffffffu, nm * bits);
                // lane t < 10: total length under table t+1; strictly shortest wins, ties -> lowest ID (Huffman.py:300-307)
                unsigned tot = 0xffffffu;
                if (lane < kNTables) {
                    tot = 0;
#pragma unroll
                    for (int q = 0; q < WPC; q++) tot += sTot[ch * WPC + q][lane];
                }
                const unsigned bk = __reduce_min_sync(0xffffffffu, (tot << 4) | (unsigned)lane);
                const unsigned best = bk >> 4;
                const int bestID = (int)(bk & 15u) + 1;
                bitDeposit += (long long)origin - ((long long)best + nMant + ec.nTableIDBits);   // codec.py:118-120
                const long long nbits = (long long)ec.fixedBits + nMant + best;                   // pacfile.py:291-312
                const unsigned nby = (unsigned)((nbits + 7) / 8);                                 // :315-316
                if (lane == 0) {
                    a.tableID[wc] = (uint8_t)bestID;
                    a.nbytes[wc] = nby;
                    a.chunkOff[wc] = outOff;
                }
                outOff += 4 + (long long)nby;
            }
            if (lane == 0 && a.trExtra) { a.trExtra[w] = extraBits; a.trDeposit[w] = bitDeposit; }
        }
    }
    if (warp == 0 && lane == 0) {
        a.state[s].extraBits = extraBits;
        a.state[s].bitDeposit = bitDeposit;
        a.state[s].outOffset = outOff;
        if (a.tileEnd) a.tileEnd[s] = outOff;
    }
}

}  // namespace pac
