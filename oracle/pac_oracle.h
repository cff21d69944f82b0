/*
 * pac_oracle.h -- CPU restatement of the reference codec's per-block hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (perceptual-audio-codec_b200/)
 * may include, link, import or execute this.  Callers allowed: tests/, __graft_entry__.smoke(),
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * Every function cites the reference file:line (relative to /root/reference/) it follows.
 * All arithmetic is IEEE double in the reference's own operation order; the FFT is a plain
 * radix-2 (the reference calls numpy's pocketfft -- outputs agree to ~1e-15 relative, and the
 * committed goldens pin the end-to-end bytes; see DESIGN.md "Oracle").
 */
#ifndef PAC_ORACLE_H
#define PAC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_BANDS 32
#define ORC_NTABLES 10

typedef struct {
    int32_t sampleRate;           /* pcmfile.py:44-45 (int from struct.unpack) */
    int32_t nChannels;            /* must be 2: codec.py:46-47, psychoac.py:477 */
    int32_t nMDCTLines;           /* pacfile.py:452 */
    int32_t nScaleBits;           /* pacfile.py:453 */
    int32_t nMantSizeBits;        /* pacfile.py:454 */
    int32_t nTableIDBits;         /* pacfile.py:457 */
    double  targetBitsPerSample;  /* pacfile.py:455 */
    int32_t window;               /* 0: SineWindow (codec.py:59-60,239-240 as HEAD has them); 1: the same two call sites with
                                   * window.KBDWindow (window.py:56-78) -- which returns a COPY, so the psychoacoustic model then
                                   * sees the un-windowed block (no Q1 aliasing) */
    int32_t reserved;
} OrcParams;

/* Flattened huffmanTables.pickle (Huffman.py:138-153): for table t (ID t+1), magnitudes
 * 0..nkeys[t]-1 have code value code[off[t]+v] of length len[off[t]+v] bits; len 0 = key absent
 * (=> escape, Huffman.py:292-298).  esc_code/esc_len = encodingTable[-1]. */
typedef struct {
    int32_t nkeys[ORC_NTABLES];
    int32_t off[ORC_NTABLES];
    const uint32_t *code;
    const uint8_t *len;
    uint32_t esc_code[ORC_NTABLES];
    int32_t esc_len[ORC_NTABLES];
} OrcHuff;

/* per-block trace of one stream's encode (any pointer may be NULL) */
typedef struct {
    int32_t *lrms;      /* [nBlocks] 25-bit mask, bit b = LRMS[b]           codec.py:96-102 */
    int32_t *oscale;    /* [nBlocks][2]                                     codec.py:245 */
    int32_t *ba;        /* [nBlocks][2][nBands]                             codec.py:258 */
    int32_t *sf;        /* [nBlocks][2][nBands]                             codec.py:274 */
    int32_t *tableID;   /* [nBlocks][2]                                     Huffman.py:309 */
    int32_t *nbytes;    /* [nBlocks][2] chunk payload bytes                 pacfile.py:291-317 */
    int64_t *extraBits; /* [nBlocks] cp.extraBits after the block           codec.py:229,260 */
    int64_t *bitDeposit;/* [nBlocks] huffman.bitDeposit after the block     codec.py:120 */
    double  *smr;       /* [nBlocks][2][nBands]                             psychoac.py:662-682 */
    double  *lines;     /* [nBlocks][2][nMDCTLines] LRMS-selected lines     psychoac.py:663-680 */
    int32_t *mant;      /* [nBlocks][2][nMDCTLines] signed codes at line positions (0 where ba==0) */
} OrcTrace;

/* ---- static layout ---- */
int  orc_band_layout(int nMDCTLines, int sampleRate, int32_t *nLines /*[25]*/);   /* psychoac.py:124-156 */

/* ---- L2 kernels ---- */
void orc_sine_window(double *x, int N);                                   /* window.py:27-39  (in place) */
void orc_hann_window(double *x, int N);                                   /* window.py:41-53  (in place) */
void orc_kbd_window(const double *x, double *out, int N, double alpha);   /* window.py:56-78  (copy) */
void orc_fft(double *re, double *im, int N, int inverse);                 /* np.fft.fft / ifft */
void orc_mdct(const double *x, int a, int b, double *X /*[(a+b)/2]*/);    /* mdct.py:49-71 */
void orc_imdct(const double *X, int a, int b, double *x /*[a+b]*/);       /* mdct.py:73-88 */
double orc_spl(double intensity);                                         /* psychoac.py:15-35 */
double orc_intensity(double spl);                                         /* psychoac.py:37-42 */
double orc_thresh(double f);                                              /* psychoac.py:44-54 */
double orc_bark(double f);                                                /* psychoac.py:56-64 */
int  orc_quantize_uniform(double x, int nBits);                           /* quantize.py:40-64 */
void orc_vquantize_uniform(const double *x, int n, int nBits, uint64_t *q);       /* quantize.py:91-117 */
void orc_vdequantize_uniform(const uint64_t *q, int n, int nBits, double *x);     /* quantize.py:120-145 */
int  orc_scale_factor(double x, int nScaleBits, int nMantBits);           /* quantize.py:148-177 */
void orc_vmantissa(const double *x, int n, int scale, int nScaleBits, int nMantBits, uint64_t *m); /* quantize.py:315-342 */
void orc_vdequantize(int scale, const int64_t *m, int n, int nScaleBits, int nMantBits, double *x); /* quantize.py:345-376 */
int  orc_bitalloc(double bitBudget, int64_t extraBits, int maxMantBits, int nBands, const int32_t *nLines,
                  const double *SMR, const int32_t *LRMS, int32_t *bits, int64_t *bitDifference); /* bitalloc.py:129-184 */

/* bitalloc.py:22-125: mode 0 BitAllocUniform, 1 BitAllocConstSNR (level = peakSPL per band), 2 BitAllocConstMNR (level = SMR);
 * returns -1 where the reference's loop would never terminate */
int  orc_bitalloc_alt(int mode, double bitBudget, int maxMantBits, int nBands, const int32_t *nLines, const double *level,
                      int32_t *bits);

/* masked threshold of one (already sine-windowed) time block; applies the Hann window IN PLACE
 * exactly as the reference does.  psychoac.py:409-456 */
void orc_calc_bthr(double *data, int N, int nMDCTLines, int sampleRate, int noDrop, double *thr);
/* mono path, psychoac.py:215-318 */
void orc_calc_smrs(double *data, int N, const double *mdct, int nMDCTLines, int mdctScale, int sampleRate,
                   const int32_t *nLines, int nBands, double *smr, double *thr /*may be NULL*/);
/* stereo path, psychoac.py:506-682.  data[2][N] = sine-windowed time samples (modified in place),
 * mdct[2][nMDCTLines] = scaled lines.  bthr6 (may be NULL) receives L,R,M,S,M',S' curves. */
void orc_stereo_smr(double *data0, double *data1, int N, const double *mdct0, const double *mdct1,
                    int nMDCTLines, const int32_t *scale, int sampleRate, const int32_t *nLines, int nBands,
                    const int32_t *LRMS, double *smr /*[2][nBands]*/, double *lines /*[2][nMDCTLines]*/,
                    double *bthr6 /*[6][nMDCTLines]*/);
/* codec.py:96-102 on raw (unwindowed) 2N-sample blocks */
void orc_lrms(const double *l, const double *r, int N, const int32_t *nLines, int nBands, int32_t *LRMS);

/* ---- whole streams ---- */
/* pcm: interleaved int16 [nSamples][2].  Returns bytes written (header + chunks), or <0:
 * -1 bad params, -2 out buffer too small.  pacfile.py:231-366, codec.py:83-281 */
int64_t orc_encode_stream(const OrcParams *p, const OrcHuff *h, const int16_t *pcm, int64_t nSamples,
                          uint8_t *out, int64_t cap, OrcTrace *trace, int64_t *final_state /*[2] dep,extra*/);
int64_t orc_encoded_blocks(int64_t nSamples, int nMDCTLines);
/* pac -> interleaved int16 PCM as the reference __main__ Decode pass writes it (first block dropped,
 * overlap tail emitted).  Returns samples per channel written, or <0 on error.
 * pacfile.py:123-229, codec.py:25-65, pcmfile.py:118-147 */
int64_t orc_decode_stream_w(const OrcHuff *h, const uint8_t *pac, int64_t nbytes, int16_t *pcm, int64_t capSamples,
                            OrcParams *hdr_out, int64_t *numSamplesHdr, int window /* as OrcParams.window: the container does not store it */);
int64_t orc_decode_stream(const OrcHuff *h, const uint8_t *pac, int64_t nbytes, int16_t *pcm, int64_t capSamples,
                          OrcParams *hdr_out, int64_t *numSamplesHdr);
/* encode many equal-length streams on `nthreads` host threads (CPU baseline).  pcm [S][nSamples][2],
 * out [S][cap], outBytes[S].  Returns 0 or the first negative status. */
int orc_encode_batch(const OrcParams *p, const OrcHuff *h, const int16_t *pcm, int64_t nSamples, int S,
                     uint8_t *out, int64_t cap, int64_t *outBytes, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
