/*
 * flo_oracle.h -- CPU oracle for flo's lossless ALPC encode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product (flo_b200/) never links,
 * imports or falls back to this code.
 *
 * Every function is a plain-C restatement of the reference algorithm and
 * cites the reference file:line it follows (paths relative to the
 * reference checkout, e.g. libflo/src/lossless/encoder.rs).
 *
 * Parity pin: see oracle/README.md -- the restatement is pinned against the
 * reference's shipped Examples/ *.flo bitstreams (tests/golden/) for raw,
 * fixed orders 0-4, LPC order 5 and the container; everything else is
 * restatement-only.
 */
#ifndef FLO_ORACLE_H
#define FLO_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- primitives (known-answer tests) ---------------------------------- */
int32_t  flo_ref_f32_to_i32(float x);                       /* core/audio_constants.rs:18-20 */
float    flo_ref_i32_to_f32(int32_t v);                     /* core/audio_constants.rs:24-26 */
uint32_t flo_ref_crc32(const uint8_t *data, size_t len);    /* core/crc32.rs:23-30 */
uint8_t  flo_ref_estimate_rice_parameter_i32(const int32_t *r, size_t n); /* core/rice.rs:29-69 */
/* returns byte length; *out is malloc'd (free with flo_ref_free) */
size_t   flo_ref_rice_encode_i32(const int32_t *r, size_t n, uint8_t k, uint8_t **out); /* rice.rs:84-114 */
void     flo_ref_rice_decode_i32(const uint8_t *enc, size_t enc_len, uint8_t k,
                                 size_t target_len, int32_t *out);        /* rice.rs:123-159 */
void     flo_ref_fixed_predictor_residuals(const int32_t *s, size_t n, int order, int32_t *out); /* lpc.rs:301-359 */
void     flo_ref_autocorr_int(const int32_t *s, size_t n, int order, int64_t *out); /* lpc.rs:213-221 */
/* returns 1 = Some, 0 = None */
int      flo_ref_levinson_durbin_int(const int64_t *ac, int order, int32_t *coeffs, uint8_t *shift); /* lpc.rs:225-276 */
void     flo_ref_calc_residuals_int(const int32_t *s, size_t n, const int32_t *coeffs,
                                    uint8_t shift, int order, int32_t *out); /* lpc.rs:279-298 */

/* ---- the encode path --------------------------------------------------- */
/* Encoder::new(sr, ch, bits).with_compression(level).encode(samples, meta)
 * (lossless/encoder.rs:17-45).  0 = ok; -1 = the reference would panic
 * (channels == 0 or sample_rate == 0).  *out malloc'd. */
int flo_ref_encode(const float *samples, size_t n_interleaved, uint32_t sample_rate,
                   uint8_t channels, uint8_t bit_depth, uint8_t level,
                   const uint8_t *meta, size_t meta_len, uint8_t **out, size_t *out_len);

/* reflo ingest arm for 16-bit PCM (reflo/src/audio.rs:247-254: s * 1/32768)
 * followed by flo_ref_encode. */
int flo_ref_encode_pcm16(const int16_t *pcm, size_t n_interleaved, uint32_t sample_rate,
                         uint8_t channels, uint8_t bit_depth, uint8_t level,
                         const uint8_t *meta, size_t meta_len, uint8_t **out, size_t *out_len);

/* Enter at the i32 boundary (after f32_to_i32/deinterleave/mid-side):
 * encode one frame from coded-domain channels exactly as encode_frame does
 * from encoder.rs:102-127, with the mid/side flag given by the caller.
 * Returns the serialised frame bytes (writer.rs:236-301). */
int flo_ref_encode_frame_i32(const int32_t *const *ch, const size_t *ch_len, uint8_t channels,
                             uint32_t frame_samples, uint8_t flags, uint8_t level,
                             uint8_t **out, size_t *out_len);

/* Per-candidate report for one channel (encoder.rs:173-217): fills up to 14
 * entries in candidate order raw, fixed 0..4, lpc 5..P.  size = -1 when the
 * candidate is None/skipped. Returns the number of entries. */
typedef struct {
    int32_t kind;      /* 0 raw, 1 fixed, 2 lpc */
    int32_t order;
    int32_t k;
    int64_t size;      /* encoded.len(), or -1 */
} flo_ref_candidate;
int flo_ref_channel_candidates(const int32_t *s, size_t n, uint8_t level, flo_ref_candidate *outc);

/* ---- reader + decoder (reader.rs, lossless/decoder.rs) ------------------ */
typedef struct flo_ref_file flo_ref_file;
/* returns NULL on parse error (message via flo_ref_last_error) */
flo_ref_file *flo_ref_parse(const uint8_t *data, size_t len);
void     flo_ref_file_free(flo_ref_file *f);
uint32_t flo_ref_file_sample_rate(const flo_ref_file *f);
uint8_t  flo_ref_file_channels(const flo_ref_file *f);
uint8_t  flo_ref_file_bit_depth(const flo_ref_file *f);
uint8_t  flo_ref_file_level(const flo_ref_file *f);
uint64_t flo_ref_file_total_samples(const flo_ref_file *f);
uint32_t flo_ref_file_crc32(const flo_ref_file *f);
uint64_t flo_ref_file_data_offset(const flo_ref_file *f);   /* byte offset of DATA chunk in file */
uint64_t flo_ref_file_data_size(const flo_ref_file *f);
uint64_t flo_ref_file_meta_size(const flo_ref_file *f);
uint32_t flo_ref_file_num_frames(const flo_ref_file *f);
/* frame i: type, samples, flags, byte offset within DATA (from TOC), size (TOC) */
void     flo_ref_file_frame_info(const flo_ref_file *f, uint32_t i, uint8_t *type, uint32_t *samples,
                                 uint8_t *flags, uint64_t *byte_offset, uint32_t *frame_size,
                                 uint32_t *timestamp_ms);
/* channel c of frame i: order marker etc. (for reporting) */
void     flo_ref_file_channel_info(const flo_ref_file *f, uint32_t i, uint32_t c, uint32_t *n_coeffs,
                                   uint8_t *shift_bits, uint8_t *encoding, uint8_t *k,
                                   uint64_t *residual_bytes, int32_t *coeffs12);
/* decode_channel_int (decoder.rs:92-149) for every channel of frame i:
 * out is channels * frame_samples int32 (coded domain, before M/S inverse). */
int      flo_ref_file_decode_frame_coded(const flo_ref_file *f, uint32_t i, int32_t *out);
/* Decoder::decode (decoder.rs:14-73): interleaved f32; *out malloc'd */
int      flo_ref_decode(const uint8_t *data, size_t len, float **out, size_t *out_n);
/* same but stops before i32_to_f32: interleaved int32 (after M/S inverse) */
int      flo_ref_decode_i32(const uint8_t *data, size_t len, int32_t **out, size_t *out_n);

/* EBU R128 integrated loudness (flo_r128.c; libflo/src/core/ebu_r128.rs:58-102, 182-313) -- parity unpinned */
void   flo_ref_kweighting_coeffs(double sample_rate, double out[10]);
double flo_ref_r128_integrated(const float *samples, size_t n_interleaved, uint8_t channels, uint32_t sample_rate,
                               double **block_energies_out, size_t *n_blocks);

void        flo_ref_free(void *p);
const char *flo_ref_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
