/*
 * flo_b200.h -- C ABI of the B200-native lossless ALPC encoder for the flo format.
 *
 * This is the drop-in boundary for ONE path of flo-audio/flo: the lossless
 * encode path behind
 *
 *     libflo_audio::Encoder::new(sample_rate, channels, bit_depth)
 *                  .with_compression(level)
 *                  .encode(samples: &[f32], metadata: &[u8]) -> FloResult<Vec<u8>>
 *
 * (reference: libflo/src/lossless/encoder.rs:17-45, re-exported at
 * libflo/src/lib.rs:20, documented in Docs/rust-api.md:44-73, called by
 * reflo/src/lib.rs:301-305).  A Rust `Encoder` shim binds these symbols with
 * `extern "C"` (see INTEGRATION.md and rust/); the bytes returned are
 * identical to what the reference encoder returns for the same arguments.
 *
 * Plain pointers and sizes only; no torch / CUDA types in any signature.
 * There is NO CPU fallback: every entry point fails (non-zero return,
 * message via flo_last_error) when no CUDA device is usable.
 *
 * Return value of every int function: 0 = ok, non-zero = error
 * (FLO_ERR_*); the shim maps that to Err(String) with flo_last_error().
 */
#ifndef FLO_B200_H
#define FLO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLO_OK            0
#define FLO_ERR_ARG       1   /* bad argument (channels == 0 / sample_rate == 0: the reference panics, encoder.rs:48-50) */
#define FLO_ERR_CUDA      2   /* CUDA runtime / driver error, or no device */
#define FLO_ERR_NOMEM     3   /* host or device allocation failed */
#define FLO_ERR_INTERNAL  4   /* device-side consistency check failed */

/* sample formats of the batched entries */
#define FLO_FMT_F32    0      /* interleaved f32, the documented Encoder::encode input (encoder.rs:32) */
#define FLO_FMT_PCM16  1      /* interleaved i16 PCM: reflo's S16 ingest arm, s * (1/32768) (reflo/src/audio.rs:247-254), then as F32 */
#define FLO_FMT_U8     2      /* interleaved u8 PCM: reflo's U8 arm, (s - 128) / 128 (reflo/src/audio.rs:263-269), then as F32 */
#define FLO_FMT_S32    3      /* interleaved i32 PCM: reflo's S32 arm, s as f32 * (1/2147483648) (reflo/src/audio.rs:255-262), then as F32 */

typedef struct flo_ctx flo_ctx;

/* One context = one GPU + its streams and scratch arenas.  Calls on one
 * context serialise on an internal mutex; distinct contexts are independent
 * (the Rust Encoder stays Send + Sync by holding Arc<ctx>, Docs/rust-api.md:374-378). */
int  flo_ctx_create(int device, flo_ctx **out);
void flo_ctx_destroy(flo_ctx *ctx);

/* Encoder::new(sr, ch, bits).with_compression(level).encode(samples, metadata)
 * (libflo/src/lossless/encoder.rs:17-45).  `level` is clamped to 9 like
 * with_compression (encoder.rs:26-29).  *out is allocated by the library
 * (release with flo_free); inputs are borrowed for the call only. */
int flo_encode(flo_ctx *ctx, const float *samples, size_t n_interleaved,
               uint32_t sample_rate, uint8_t channels, uint8_t bit_depth, uint8_t level,
               const uint8_t *meta, size_t meta_len, uint8_t **out, size_t *out_len);

/* reflo's 16-bit ingest (reflo/src/audio.rs:247-254) fused in front of
 * flo_encode: identical bytes to flo_encode(f32(pcm) * (1/32768)). */
int flo_encode_pcm16(flo_ctx *ctx, const int16_t *pcm, size_t n_interleaved,
                     uint32_t sample_rate, uint8_t channels, uint8_t bit_depth, uint8_t level,
                     const uint8_t *meta, size_t meta_len, uint8_t **out, size_t *out_len);

/* StreamingEncoder::encode_frame_data (libflo/src/streaming/encoder.rs:216-257) for every frame of `samples`
 * (whole 1-second frames and, if the length is not a multiple, one partial frame at the end): one device pass
 * for the whole run, then the reference's per-frame layout -- frame_type, frame_samples u32, flags, and per
 * channel a u32 length + [rice_parameter][coefficients as i32 LE][residual bytes] (serialize_channel, :243-257).
 * *out holds the frames back to back, frame i = (*out)[(*frame_off)[i] .. (*frame_off)[i + 1]); *frame_off has
 * *n_frames + 1 entries.  Both are malloc'd by the library: release each with flo_free. */
int flo_stream_encode_frames(flo_ctx *ctx, const float *samples, size_t n_interleaved,
                             uint32_t sample_rate, uint8_t channels, uint8_t bit_depth, uint8_t level,
                             uint8_t **out, size_t *out_len, uint64_t **frame_off, uint32_t *n_frames);

/* Batched entry (additive; the reference has no batch call -- it is a loop of
 * Encoder::encode over tracks, reflo/src/main.rs:218-276).  Frames of all
 * tracks are encoded in one device pass.  Every track gets exactly the bytes
 * flo_encode would return for it. */
typedef struct {
    const void    *samples;        /* interleaved f32 or i16 (see `format` of the call) */
    size_t         n_interleaved;  /* number of interleaved samples (frames * channels [+ ragged tail]) */
    uint32_t       sample_rate;
    uint8_t        channels;
    uint8_t        bit_depth;      /* copied to the header only (encoder.rs:40, writer.rs:163) */
    const uint8_t *meta;           /* opaque metadata bytes appended verbatim (writer.rs:96), may be NULL */
    size_t         meta_len;
} flo_track;

typedef struct {
    uint8_t *data;                 /* library-allocated .flo file image (flo_free) */
    size_t   len;
} flo_out;

int flo_encode_batch(flo_ctx *ctx, const flo_track *tracks, size_t n_tracks, int format,
                     uint8_t level, flo_out *outs);

/* Device-resident variant: tracks[i].samples are DEVICE pointers on the
 * context's GPU (meta stays a host pointer); the concatenated file images are
 * written to the device buffer d_out (capacity >= flo_output_bound).  The
 * image of track i is d_out[offsets[i] .. offsets[i] + lens[i]); offsets/lens
 * are host arrays of n_tracks entries.  The call returns after the device
 * work finished.
 *
 * Stream ordering (both device entries, encode and decode): the kernels run on
 * the context's stream.  The context's own stream is a blocking stream, i.e. it
 * is ordered after everything issued earlier on the legacy default stream (where
 * torch's default stream runs), so inputs produced there need no extra
 * synchronisation.  Inputs produced on ANY OTHER stream must either be complete
 * (cudaStreamSynchronize / event wait) before the call, or that stream must be
 * given to the context with flo_ctx_set_stream first.  Reading half-written
 * input is not detected: the result would be a valid file of the wrong samples. */
int flo_encode_batch_device(flo_ctx *ctx, const flo_track *tracks, size_t n_tracks, int format,
                            uint8_t level, void *d_out, size_t d_out_capacity,
                            uint64_t *offsets, uint64_t *lens);

/* Worst-case total size of the file images of a batch (raw-coded frames). */
size_t flo_output_bound(const flo_track *tracks, size_t n_tracks);

/* Run the device work of this context on a caller-owned CUDA stream
 * (a cudaStream_t / CUstream passed as void*; NULL = the context's own): use the
 * stream that produces the device inputs / consumes the device outputs. */
int flo_ctx_set_stream(flo_ctx *ctx, void *cuda_stream);

/* Timing of the last batch call, measured with CUDA events on the stream the
 * kernels ran on.  ms[0] = whole device pass (first kernel .. last kernel),
 * ms[1] = frame-encode kernel, ms[2] = CRC kernels, ms[3] = setup/finalise
 * kernels, ms[4] = H2D copies, ms[5] = D2H copies.  launches = kernels
 * launched by that call. */
int flo_ctx_last_timing(flo_ctx *ctx, float ms[6], uint32_t *launches);

/* Analysis counters of the last batch call (diagnostics): [0] non-silent frames, [1] exact-size
 * re-evaluation rounds, [2] LPC candidates sized in the single pass, [3] LPC candidates that needed the
 * exact pass, [4] fixed candidates evaluated exactly, [5] candidates excluded by size bounds;
 * [8..23] SM clocks (thread 0) summed over frames: ingest, analysis, look-back, pack, whole frame,
 * pack codes / scan / emit, pass 1, Levinson, pass 2, exact rounds + selection. */
int flo_ctx_last_counters(flo_ctx *ctx, uint64_t out[24]);

/* Per-frame analysis report of the last batch call, for parity tests: for
 * global frame g and channel c (< 8), candidate j (raw, fixed 0..4, lpc
 * 5..12 -> j = 0..13): k and encoded size (-1 = candidate absent).  Must be
 * enabled before the call; costs a little device memory. */
typedef struct { int32_t k; int32_t pad; int64_t size; } flo_cand_report;
int flo_ctx_enable_report(flo_ctx *ctx, int enable);
int flo_ctx_read_report(flo_ctx *ctx, uint32_t frame, uint32_t channel, flo_cand_report out[14]);

/* ---- lossless decoder (companion of the encode path) ----
 * Decoder::decode(&self, data: &[u8]) -> FloResult<Vec<f32>> (libflo/src/lossless/decoder.rs:14-18) over
 * Reader::read (libflo/src/reader.rs:16-53): interleaved f32 samples, sample * (1/32767)
 * (core/audio_constants.rs:24-26).  Errors carry the reference's messages ("Invalid flo file: bad magic",
 * "Unexpected end of file", "Invalid TOC: too many entries", "Invalid frame: too many samples",
 * "Invalid LPC order") through flo_last_error().  Files holding transform (lossy) frames are refused. */
typedef struct {
    uint32_t sample_rate;
    uint8_t  channels, bit_depth, level, version_major;
    uint64_t total_samples;        /* header field (sample frames per channel) */
    uint64_t decoded_frames;       /* sample frames decoded = sum of frame_samples over the frames read */
    uint32_t n_frames;             /* frames read (reader.rs:116-118 may stop early) */
    uint32_t data_crc32;           /* header field; not verified (the reference decoder does not either) */
    uint64_t meta_offset, meta_size;
} flo_info;

/* *out is library-allocated (flo_free); *n_interleaved = decoded_frames * channels.  info may be NULL. */
int flo_decode(flo_ctx *ctx, const uint8_t *file, size_t len, float **out, size_t *n_interleaved, flo_info *info);

/* Device-resident variant: d_file is a DEVICE pointer to the file image (readable up to the next 16-byte
 * boundary past its end -- true for any image inside a cudaMalloc'd buffer, e.g. flo_encode_batch_device
 * output); samples are written to the device buffer d_out (capacity in floats).  When the capacity is too
 * small the call fails with FLO_ERR_ARG and *n_interleaved holds the count needed. */
int flo_decode_device(flo_ctx *ctx, const void *d_file, size_t len, float *d_out, size_t d_out_capacity,
                      size_t *n_interleaved, flo_info *info);

/* The same decode with 16-bit output: the integer samples Decoder::decode holds before its final i32 -> f32
 * conversion (lossless/decoder.rs:61-72: `sample as f32 / 32767.0`), saturated to i16 -- for callers that feed a
 * 16-bit sink (WAV writer, sound device): half the bytes leave the device.  Same errors, same info; capacity and
 * *n_interleaved count samples.  flo_decode's f32 result equals (float)i16 * (1/32767) wherever |sample| <= 32767. */
int flo_decode_i16(flo_ctx *ctx, const uint8_t *file, size_t len, int16_t **out, size_t *n_interleaved, flo_info *info);
int flo_decode_i16_device(flo_ctx *ctx, const void *d_file, size_t len, int16_t *d_out, size_t d_out_capacity,
                          size_t *n_interleaved, flo_info *info);

/* -------------------------------------------------------------------------
 * Waveform peaks of the analysis metadata libflo::encode() attaches.  Replaces
 *   core::analysis::extract_waveform_peaks(samples, channels, sample_rate, peaks_per_second) -> WaveformData
 * (libflo/src/core/analysis.rs:38-119; called from add_analysis_data_if_missing, libflo/src/lib.rs:219-241, with
 * peaks_per_second = 50): *peaks holds WaveformData.peaks (library-allocated, flo_free), already normalised to
 * the largest peak.  Every operation is an IEEE f32 max / add / divide, so the values are the reference's bit
 * for bit (two unspecified corners: the sign of a zero peak for more than two channels, and the payload of the
 * NaN an infinite largest peak leaves).  channels == 0 or sample_rate == 0 make the reference panic (capacity
 * overflow): FLO_ERR_ARG.  peaks_per_second == 0 or no samples: zero peaks.
 * The spectral fingerprint and EBU R128 parts of that metadata, and its MessagePack serialisation, are not
 * provided. */
int flo_waveform_peaks(flo_ctx *ctx, const float *samples, size_t n_interleaved, uint32_t sample_rate,
                       uint8_t channels, uint32_t peaks_per_second, float **peaks, size_t *n_peaks);
/* device-resident samples and peaks (capacity in floats; flo_waveform_peaks_count gives the count needed) */
int flo_waveform_peaks_device(flo_ctx *ctx, const float *d_samples, size_t n_interleaved, uint32_t sample_rate,
                              uint8_t channels, uint32_t peaks_per_second, float *d_peaks, size_t capacity, size_t *n_peaks);
size_t flo_waveform_peaks_count(size_t n_interleaved, uint32_t sample_rate, uint8_t channels, uint32_t peaks_per_second);

/* EBU R128 integrated loudness, the value libflo::encode() stores as loudness_profile[0].lufs (cast to f32).
 * Replaces compute_ebu_r128_loudness(samples, channels, sample_rate).integrated_lufs
 * (libflo/src/core/ebu_r128.rs:182-313, called from libflo/src/lib.rs:256-268): K-weighting per BS.1770
 * (two biquads per channel), 400 ms blocks every 100 ms, -70 LUFS absolute and -10 LU relative gates.  The
 * reference filters each channel as one serial f64 recurrence; the device filters its 100 ms hops in parallel and
 * chains their states, so *lufs agrees with the reference to ~1e-12 LU, not bit for bit.  No samples,
 * channels == 0 or no block above the absolute gate: -23.0, as in the reference.  Sample rates below ~3.4 kHz,
 * where the reference's shelf filter is unstable and its result overflows to inf, are refused (FLO_ERR_ARG).  loudness_range_lu, true_peak_dbtp
 * and sample_peak_dbfs of LoudnessMetrics are not computed (encode() does not store them). */
int flo_integrated_loudness(flo_ctx *ctx, const float *samples, size_t n_interleaved, uint32_t sample_rate,
                            uint8_t channels, double *lufs);
int flo_integrated_loudness_device(flo_ctx *ctx, const float *d_samples, size_t n_interleaved, uint32_t sample_rate,
                                   uint8_t channels, double *lufs);

/* Pinned host memory (optional).  Page-locked inputs go to the copy engine
 * directly; pageable inputs (a plain Rust slice) are staged by the library
 * through its own pinned ring with several copy threads (FLO_B200_COPY_THREADS
 * overrides their number). */
void *flo_host_alloc(size_t bytes);
void  flo_host_free(void *p);

void        flo_free(void *p);
const char *flo_last_error(void);      /* thread-local, never NULL */
const char *flo_version(void);
int         flo_device_count(void);    /* 0 when no usable CUDA device */

#ifdef __cplusplus
}
#endif
#endif /* FLO_B200_H */
