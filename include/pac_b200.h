/*
 * pac_b200.h -- C ABI of the B200-native engine for the WAK/"PAC" perceptual audio codec hot path.
 *
 * The reference (wisamreid/Perceptual-Audio-Codec) is pure Python and has no FFI; its boundary is the
 * Python call surface PACFile.WriteDataBlock/ReadDataBlock -> codec.Encode/Decode.
 * This header is what the Python shim modules in perceptual-audio-codec_b200/ (pacfile.py, codec.py,
 * psychoac.py, mdct.py, window.py, quantize.py, bitalloc.py) bind through ctypes.  Each entry point
 * names the reference interface it replaces (paths relative to /root/reference/codec/).
 *
 * Conventions
 *   - plain C types only; every pointer is a HOST pointer unless the comment says "host or device"
 *     (those are resolved with cudaPointerGetAttributes);
 *   - return value: 0 = ok, <0 = error (PAC_E_*); pac_last_error(ctx) gives the text;
 *   - one context per GPU and host thread; contexts share nothing;
 *   - there is NO CPU fallback: every function below launches sm_100a kernels or fails.
 */
#ifndef PAC_B200_H
#define PAC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAC_OK 0
#define PAC_E_ARG      (-1)   /* bad argument / unsupported parameter combination */
#define PAC_E_CUDA     (-2)   /* CUDA runtime error (text in pac_last_error) */
#define PAC_E_OVERFLOW (-3)   /* an output buffer was too small (no bytes were written past it) */
#define PAC_E_FORMAT   (-4)   /* malformed .pac input */
#define PAC_E_NODEVICE (-5)   /* no usable CUDA device */

#define PAC_PRECISION_FP64 0  /* verification mode: bit-exact .pac bytes vs the reference */
#define PAC_PRECISION_FP32 1  /* fast mode */

#define PAC_NTABLES 10
#define PAC_MAX_BANDS 32

typedef struct PacCtx PacCtx;

/* CodingParams attributes set at pacfile.py:452-457 (+ sampleRate/nChannels from pcmfile.py:44-60) */
typedef struct {
    int32_t sampleRate;
    int32_t nChannels;            /* must be 2 (codec.py:46-47, psychoac.py:477) */
    int32_t nMDCTLines;           /* 1024 (also 512/256 for the L2 entry points) */
    int32_t nScaleBits;           /* 4 */
    int32_t nMantSizeBits;        /* 4 */
    int32_t nTableIDBits;         /* 4 */
    double  targetBitsPerSample;  /* 2.27 */
    int32_t window;               /* PAC_WINDOW_SINE (what HEAD's codec calls, codec.py:59-60,239-240) or PAC_WINDOW_KBD: the same two
                                   * call sites with window.KBDWindow (window.py:56-78, alpha = 4) in place of SineWindow.  KBDWindow
                                   * returns a COPY, so -- unlike the in-place SineWindow -- the psychoacoustic model
                                   * then sees the un-windowed block. */
    int32_t reserved;
} PacParams;
#define PAC_WINDOW_SINE 0
#define PAC_WINDOW_KBD  1

/* huffmanTables.pickle flattened by the host shim (Huffman.py:138-153, 256-262): for table ID t+1,
 * magnitude v < nkeys[t] has code value code[off[t]+v] of len[off[t]+v] bits; len 0 = key absent
 * (escape, Huffman.py:292-298).  esc_* = encodingTable[-1]. */
typedef struct {
    int32_t nkeys[PAC_NTABLES];
    int32_t off[PAC_NTABLES];
    const uint32_t *code;
    const uint8_t *len;
    uint32_t esc_code[PAC_NTABLES];
    int32_t esc_len[PAC_NTABLES];
} PacHuffTables;

/* Optional per-block taps of pac_encode_batch (parity tests; any pointer may be NULL).
 * B = max over streams of pac_num_blocks(nSamples[s]); arrays are [S][B]-major.
 * Floating taps are always double on the host side (fp32 mode widens). */
typedef struct {
    int32_t *lrms;        /* [S][B]           bit b = LRMS[b]                      codec.py:96-102 */
    int32_t *oscale;      /* [S][B][2]        overallScaleFactor                   codec.py:245 */
    double  *smr;         /* [S][B][2][nBands]                                     psychoac.py:662-682 */
    double  *lines;       /* [S][B][2][nMDCTLines] LRMS-selected, scaled lines     psychoac.py:663-680 */
    int32_t *ba;          /* [S][B][2][nBands] bitAlloc                            codec.py:258 */
    int32_t *sf;          /* [S][B][2][nBands] scaleFactor                         codec.py:274 */
    int32_t *tableID;     /* [S][B][2]                                             Huffman.py:309 */
    int32_t *nbytes;      /* [S][B][2]        chunk payload bytes                  pacfile.py:291-317 */
    int64_t *extraBits;   /* [S][B]           cp.extraBits after the block         codec.py:229,260 */
    int64_t *bitDeposit;  /* [S][B]           huffman.bitDeposit after the block   codec.py:120 */
    int32_t *mant;        /* [S][B][2][nMDCTLines] signed mantissa codes at line positions (0 where bitAlloc == 0)  codec.py:276-277 */
} PacTrace;

/* ------------------------------------------------------------------ context */
/* replaces: Huffman() + the CodingParams set-up in pacfile.py:394,446-471 */
int  pac_ctx_create(int device, int precision, const PacParams *params, const PacHuffTables *tables, PacCtx **out);
void pac_ctx_destroy(PacCtx *ctx);
const char *pac_last_error(PacCtx *ctx);          /* ctx may be NULL: error of the last failed pac_ctx_create */
const char *pac_version(void);
/* psychoac.py:124-156 + ScaleFactorBands :193-213 */
int  pac_band_layout(PacCtx *ctx, int32_t *nLines /*[PAC_MAX_BANDS]*/, int32_t *nBands);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t pac_launch_count(PacCtx *ctx);

/* run this context's kernels and copies on the caller's CUDA stream (a cudaStream_t), so that the caller's CUDA events
 * bracket the work and work the caller has queued on that stream (e.g. the kernel that produced a device-resident
 * input) is ordered before the library's.  A NULL handle means what it means to CUDA: the legacy default stream
 * (torch's default stream reports handle 0) -- it is mapped to cudaStreamLegacy, NOT to a private stream.
 * PAC_STREAM_OWN restores the context's own non-blocking stream (the state after pac_ctx_create). */
#define PAC_STREAM_OWN ((void *)(intptr_t)-1)
int pac_set_stream(PacCtx *ctx, void *stream);

/* per-kernel device time, measured with CUDA events on the stream the kernels are launched on (bench.py's roofline).
 * kinds index ms[] / count[]; timing adds two event records per launch and is off by default. */
#define PAC_K_ANALYSIS 0   /* window + MDCT + M/S decision + SMR (analysis.cuh) */
#define PAC_K_SCAN     1   /* reservoir scan: BitAlloc, scale factors, Huffman table search (scan.cuh) */
#define PAC_K_PACK     2   /* quantise + Huffman code + bit pack (pack.cuh) */
#define PAC_K_INDEX    3   /* decoder: chunk chain walk */
#define PAC_K_UNPACK   4   /* decoder: bit unpack + Huffman decode + dequantise */
#define PAC_K_SYNTH    5   /* decoder: M/S recombine + IMDCT + window + overlap-add + PCM */
#define PAC_K_MDCT     6   /* window + MDCT + overall scale by itself (pac_mdct_batch; the MDCT-only instantiation of the analysis kernel) */
#define PAC_NKINDS     8
int pac_timing_enable(PacCtx *ctx, int on);                       /* also resets the accumulators */
int pac_timing_get(PacCtx *ctx, double *ms /*[PAC_NKINDS]*/, int64_t *count /*[PAC_NKINDS]*/);

/* Page-locked, device-mapped host memory for the whole-stream calls' host buffers (cudaHostAlloc): PCM slabs read straight from
 * WAV files are copied by DMA without a pageable bounce, and a pinned `out` is written by the pack kernel in place (see
 * pac_encode_batch).  NULL on failure.  Not tied to a context; free with pac_pinned_free. */
void *pac_pinned_alloc(size_t nbytes);
void  pac_pinned_free(void *p);
/* measured host -> device copy bandwidth (GB/s) of a host buffer on `device`: for callers that place pinned memory by measurement
 * when the OS does not tell which NUMA node a GPU hangs off (bench.py's e2e leg) */
int   pac_h2d_bandwidth(const void *host, size_t nbytes, int device, double *gbs);

/* ------------------------------------------------------------------ whole streams (the hot path) */
/* ceil(n/nMDCTLines)+1: pcmfile.py:66-82 + the flush block of pacfile.py:355-365 */
int64_t pac_num_blocks(PacCtx *ctx, int64_t nSamples);
/* safe per-stream output capacity for pac_encode_batch */
int64_t pac_encode_bound(PacCtx *ctx, int64_t nSamples);

/* replaces the Encode pass of pacfile.py:430-499 for S independent streams:
 *   PCMFile.ReadDataBlock (pcmfile.py:66-100) -> PACFile.WriteDataBlock (pacfile.py:273-353) -> codec.Encode
 *   (codec.py:83-129) -> EncodeDualChannel (:212-281) -> getStereoMaskThreshold (psychoac.py:506-682) -> BitAlloc
 *   (bitalloc.py:129-184) -> ScaleFactor/vMantissa (quantize.py:148-177,315-342) -> Huffman.encodeData
 *   (Huffman.py:274-309) -> PackedBits.WriteBits (bitpack.py:36-101), plus WriteFileHeader (pacfile.py:231-271)
 *   and the flush block of Close (pacfile.py:355-366).
 * pcm: interleaved int16 [S][strideSamples][2], host or device.  nSamples[s] <= strideSamples (host).
 * out: [S][cap] bytes, host or device; outBytes[s] (host) = bytes of stream s's complete .pac file image.  Only
 *      out[s][0 .. outBytes[s]) is defined afterwards: the rest of a row is left as it was.
 * Host buffers: a pcm batch that fits one device staging buffer (up to 64 GB or half of the free device memory) is copied in slab by
 *      slab ahead of the kernels, larger ones in double-buffered stream groups.  A PINNED (page-locked, device-mapped:
 *      pac_pinned_alloc) out is filled while the kernels run -- only outBytes[s] bytes of an image cross PCIe, there is no copy-back
 *      phase; a pageable out is staged on the device and copied back group by group.
 * finalState (host, may be NULL): [S][2] = (huffman.bitDeposit, cp.extraBits) at end of stream. */
int pac_encode_batch(PacCtx *ctx, const int16_t *pcm, int64_t strideSamples, const int64_t *nSamples, int S,
                     uint8_t *out, int64_t cap, int64_t *outBytes, int64_t *finalState, const PacTrace *trace);

/* replaces the Decode pass of pacfile.py:430-499: PACFile.ReadFileHeader/ReadDataBlock (pacfile.py:123-229),
 * Huffman.decodeData (Huffman.py:321-344), codec.Decode (codec.py:25-65), overlap-add, first block dropped
 * (pacfile.py:485-487), tail emitted (:171-176), PCMFile.WriteDataBlock quantisation (pcmfile.py:118-147).
 * pac: concatenated file images, stream s = pac[pacOff[s] .. pacOff[s+1]) (pacOff host; pac host or device).  A DEVICE
 *      buffer is read in aligned 16-byte granules, so it must sit in an allocation that extends to the next multiple of
 *      16 bytes past the last image (any cudaMalloc'd buffer does); host images are staged by the library.
 * pcm: interleaved int16 [S][strideSamples][2], host or device; nSamplesOut[s] (host) = samples/channel written;
 * hdrNumSamples/hdrSampleRate (host, may be NULL) = header fields (the WAV header uses them, pcmfile.py:103-116). */
int pac_decode_batch(PacCtx *ctx, const uint8_t *pac, const int64_t *pacOff, int S, int16_t *pcm,
                     int64_t strideSamples, int64_t *nSamplesOut, int64_t *hdrNumSamples, int32_t *hdrSampleRate);
/* same, for images that are not back to back: stream s = pac[pacBeg[s] .. pacBeg[s]+pacLen[s]) -- e.g. the [S][cap] output
 * of pac_encode_batch decoded in place on the device (pacBeg[s] = s*cap, pacLen[s] = outBytes[s]). */
int pac_decode_batch_strided(PacCtx *ctx, const uint8_t *pac, const int64_t *pacBeg, const int64_t *pacLen, int S, int16_t *pcm,
                             int64_t strideSamples, int64_t *nSamplesOut, int64_t *hdrNumSamples, int32_t *hdrSampleRate);
/* The window + MDCT stage of pac_encode_batch by itself, over the same (stream, block) tiling: PCMFile.ReadDataBlock's
 * int16 -> fraction (pcmfile.py:91-98), SineWindow (window.py:27-39), MDCT (mdct.py:49-71), overall scale (codec.py:237-246).
 * pcm / strideSamples / nSamples as pac_encode_batch.  lines (host, may be NULL): [S][B][2][nMDCTLines] the scaled L/R lines
 * (2^overallScale * MDCT), oscale (host, may be NULL): [S][B][2], B = max pac_num_blocks; with both NULL the results stay in the
 * library's tile workspace (stage timing).  deviceMs (may be NULL): CUDA-event time of the kernel launches only. */
int pac_mdct_batch(PacCtx *ctx, const int16_t *pcm, int64_t strideSamples, const int64_t *nSamples, int S, double *lines,
                   int32_t *oscale, double *deviceMs);
/* The whole analysis stage of pac_encode_batch (window + MDCT + M/S decision + stereo SMR: the k_analysis kernel) over the same
 * tiling, by itself -- not overlapped with the scan/pack kernels as inside pac_encode_batch -- results left in the tile
 * workspace.  For stage timing: SMR stage = this - pac_mdct_batch.  deviceMs: CUDA-event time of the kernel launches. */
int pac_analysis_batch(PacCtx *ctx, const int16_t *pcm, int64_t strideSamples, const int64_t *nSamples, int S, double *deviceMs);
/* upper bound of samples/channel a .pac image of nbytes can decode to */
int64_t pac_decode_bound(PacCtx *ctx, int64_t nbytes);

/* ------------------------------------------------------------------ per-block API (codec.Encode / codec.Decode) */
/* Mutable per-stream state the reference keeps on CodingParams / Huffman objects */
typedef struct {
    int64_t extraBits;    /* cp.extraBits      pacfile.py:269, codec.py:229,260 */
    int64_t bitDeposit;   /* huffman.bitDeposit Huffman.py:262,353-371 */
} PacStreamState;

/* codec.Encode(data, codingParams, huffman) (codec.py:83-129) on nblk independent (state, block) pairs.
 * data [nblk][2][2*nMDCTLines] signed fractions (prior|current, NOT windowed; unlike the reference the
 * caller's array is not modified).  Outputs (host):
 *   scaleFactor,bitAlloc [nblk][2][nBands]; mant [nblk][2][nMDCTLines] signed mantissa codes at their line
 *   positions (0 where bitAlloc==0); tableID, overallScale [nblk][2]; lrms [nblk] bit mask;
 *   chunk (may be NULL) [nblk][2][chunkCap] packed payload as WriteDataBlock writes it (pacfile.py:319-351),
 *   chunkBytes [nblk][2].  state[nblk] is updated in place. */
int pac_encode_blocks(PacCtx *ctx, const double *data, int nblk, PacStreamState *state,
                      int32_t *scaleFactor, int32_t *bitAlloc, int32_t *mant, int32_t *tableID,
                      int32_t *overallScale, int32_t *lrms, uint8_t *chunk, int64_t chunkCap, int32_t *chunkBytes);

/* codec.Decode(scaleFactor, bitAlloc, mantissa, overallScaleFactor, codingParams, LRMS) (codec.py:25-65).
 * mant [nblk][2][nMDCTLines] signed codes at line positions; out [nblk][2][2*nMDCTLines] windowed IMDCT output
 * (pre overlap-add), including the decoder's M/S aliasing quirk (codec.py:46-56). */
int pac_decode_blocks(PacCtx *ctx, const int32_t *scaleFactor, const int32_t *bitAlloc, const int32_t *mant,
                      const int32_t *overallScale, const int32_t *lrms, int nblk, double *out);
/* parse nblk (2-channel) chunks as ReadDataBlock does (pacfile.py:167-217, Huffman.py:321-344).
 * chunks: [nblk][2][chunkCap] payloads, chunkBytes [nblk][2]. */
int pac_unpack_blocks(PacCtx *ctx, const uint8_t *chunks, int64_t chunkCap, const int32_t *chunkBytes, int nblk,
                      int32_t *scaleFactor, int32_t *bitAlloc, int32_t *mant, int32_t *overallScale, int32_t *lrms,
                      int32_t *tableID);

/* ------------------------------------------------------------------ L2 entry points (reference function names) */
/* window.py:27-39 SineWindow (kind 0), :41-53 HanningWindow (1), :56-78 KBDWindow alpha=4 (2); x [n][N] in place */
int pac_window(PacCtx *ctx, int kind, double *x, int n, int N);
/* mdct.py:49-71 MDCT(data, N/2, N/2): x [n][N] -> X [n][N/2];  mdct.py:73-88 IMDCT: X [n][N/2] -> x [n][N] */
int pac_mdct(PacCtx *ctx, const double *x, int n, int N, double *X);
int pac_imdct(PacCtx *ctx, const double *X, int n, int N, double *x);
/* the analysis stage on nblk raw blocks: codec.py:96-102 (LRMS) + :237-246 (window, MDCT, overall scale) +
 * psychoac.getStereoMaskThreshold (psychoac.py:506-682).  data [nblk][2][N] raw signed fractions.
 * mdct [nblk][2][N/2] scaled L/R lines, bthr [nblk][6][N/2] (L,R,M,S,M',S' thresholds, dB), smr [nblk][2][nBands],
 * lines [nblk][2][N/2] LRMS-selected lines; any output may be NULL. */
int pac_analysis(PacCtx *ctx, const double *data, int nblk, int32_t *lrms, int32_t *oscale, double *mdct,
                 double *bthr, double *smr, double *lines);
/* psychoac.CalcSMRs (psychoac.py:253-318), mono: data [n][N] time samples (window NOT yet applied; the reference
 * applies Hann itself), mdct [n][N/2] lines scaled by 2^scale, -> smr [n][nBands] */
int pac_calc_smrs(PacCtx *ctx, const double *data, const double *mdct, int n, int scale, double *smr);
/* psychoac.getMaskedThreshold (psychoac.py:215-251) / calcBTHR (:409-456) on n mono blocks: data [n][N] (the kernel
 * applies the Hann window, :225/:428), noDrop as calcBTHR's flag -> thr [n][N/2] in dB */
int pac_masked_threshold(PacCtx *ctx, const double *data, int n, int noDrop, double *thr);
/* Huffman.encodeData's table search (Huffman.py:284-308) on n caller-supplied unsigned mantissas with their per-symbol
 * bit allocation: tableID (1..10) and the total code bits under each table, totals [PAC_NTABLES] */
int pac_huffman_select(PacCtx *ctx, const uint32_t *mag, const int32_t *ba, int n, int32_t *tableID, int64_t *totals);
/* bitalloc.BitAlloc (bitalloc.py:129-184) on n independent problems */
int pac_bitalloc(PacCtx *ctx, int n, const double *bitBudget, const int64_t *extraBits, int maxMantBits,
                 const double *smr /*[n][nBands]*/, const int32_t *lrms /*[n] masks*/, int32_t *bits /*[n][nBands]*/,
                 int64_t *bitDifference /*[n]*/);
/* The allocators HEAD's codec does not call, kept for API parity: mode 0 bitalloc.BitAllocUniform (bitalloc.py:22-57, level may be
 * NULL), 1 BitAllocConstSNR (:60-91, level[n][nBands] = peakSPL replicated per band), 2 BitAllocConstMNR (:94-125, level = SMR), on n
 * independent problems.  Returns PAC_E_ARG where the reference's `while remaining_bits > 0` would never terminate (bits left, no band
 * able to take one). */
int pac_bitalloc_alt(PacCtx *ctx, int mode, int n, const double *bitBudget, int maxMantBits, const double *level /*[n][nBands]*/,
                     int32_t *bits /*[n][nBands]*/);
/* Histogram.generateStatistics (Huffman.py:71-83), the data-parallel half of HuffmanTrainer.countFreq (:184-185): occurrences of every
 * unsigned mantissa code < nbins among codes[0..n) (host or device pointer) and the position base+i of each code's first occurrence
 * (-1 where absent; the reference's dict keeps first-insertion order, which breaks frequency ties in makeHuffmanNodeQueue :93-108).
 * Codes >= nbins are ignored.  The tree itself (:221-246) is a few hundred nodes of host work and stays in the Python shim. */
int pac_histogram(PacCtx *ctx, const uint32_t *codes, int64_t n, int64_t base, int nbins, int64_t *counts /*[nbins]*/,
                  int64_t *first /*[nbins]*/);
/* quantize.py: ScaleFactor :148-177 (n scalars), vQuantizeUniform :91-117, vDequantizeUniform :120-145,
 * vMantissa :315-342, vDequantize :345-376 */
int pac_scale_factor(PacCtx *ctx, const double *x, int n, int nScaleBits, int nMantBits, int32_t *scale);
int pac_vquantize_uniform(PacCtx *ctx, const double *x, int n, int nBits, uint64_t *q);
int pac_vdequantize_uniform(PacCtx *ctx, const uint64_t *q, int n, int nBits, double *x);
int pac_vmantissa(PacCtx *ctx, const double *x, int n, int scale, int nScaleBits, int nMantBits, uint64_t *m);
int pac_vdequantize(PacCtx *ctx, int scale, const int64_t *m, int n, int nScaleBits, int nMantBits, double *x);

#ifdef __cplusplus
}
#endif
#endif
