// flo_internal.h -- structures shared by the kernels (flo_kernels.cu) and the
// host side of the C ABI (flo_api.cu).  Not part of the public interface.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/flo_b200.h"

namespace flo {

// The frame-encode kernel is built in three variants (threads per CTA x CTAs per SM): 512 x 1 keeps a whole
// 44.1/48 kHz stereo frame in shared memory; 256 x 2 and 128 x 4 run several smaller frames per SM, which
// hides the per-frame serial sections (Levinson, look-back, barriers) of small frames behind each other.
#ifndef FLO_CH
#define FLO_CH 16
#endif
constexpr int CH = FLO_CH;            // samples per thread chunk (8 or 16)
constexpr int RING_WORDS = 4096;      // bit-packer staging ring (16 KB)
constexpr int MAXORD = 12;
constexpr int NCAND = 14;             // raw, fixed 0..4, lpc 5..12
constexpr int REPORT_CH = 8;          // channels per frame covered by the parity report
constexpr uint32_t FILE_HDR = 70;     // magic(4) + header(66), writer.rs:132-191
#ifndef FLO_CRC_SEG
#define FLO_CRC_SEG 65536
#endif
constexpr uint32_t CRC_SEG = FLO_CRC_SEG;   // bytes of DATA per CRC segment
constexpr int CRC_NT = 256;

// One track of the batch (device copy).
struct TrackDev {
    const void *samples;              // device pointer, interleaved f32 or i16
    unsigned long long n_inter;       // interleaved samples
    unsigned long long static_off;    // bytes of all earlier tracks' header+TOC+meta (file start = static_off + excl[first_frame])
    unsigned long long meta_off;      // into the metadata arena
    unsigned long long meta_len;
    uint32_t sample_rate;
    uint32_t channels;
    uint32_t bit_depth;
    uint32_t first_frame;             // global index of this track's first frame
    uint32_t n_frames;
    uint32_t first_seg;               // first CRC segment slot of this track
};

// Winner of the predictor search for one channel of one frame.
struct ChanResult {
    int32_t kind;                     // 0 raw, 1 fixed, 2 lpc, 3 empty
    int32_t order;
    int32_t k;
    uint32_t nbytes;                  // residual payload bytes
    int32_t coef[MAXORD];
    int32_t shift;                    // LPC shift_bits (lpc.rs:266-268; always 15 in practice)
    int32_t pad[3];
    unsigned long long regbits[16];   // payload bits of each packer region (one region per warp of the channel)
};

struct EncodeParams {
    const TrackDev *tracks;
    const uint2 *frames;              // per global frame: {track, frame index in track}
    uint32_t n_frames;
    uint32_t frame_begin, frame_end;  // this launch encodes global frames [frame_begin, frame_end)
    int format;                       // FLO_FMT_*
    int level;                        // 0..9
    uint8_t *out;                     // output arena
    unsigned long long *status;       // decoupled look-back words, one per frame (zeroed)
    uint32_t *ticket;                 // dynamic frame counter of this launch (zeroed): next frame = frame_begin + ticket++
    unsigned long long *frame_excl;   // out: exclusive prefix of frame sizes (global)
    uint32_t *frame_size;             // out
    int16_t *plane_scratch;           // per-CTA global sample planes for frames too large for shared memory
    unsigned long long plane_scratch_elems;  // int16 elements per CTA
    ChanResult *cres;                 // per-CTA channel results, 256 entries per CTA
    flo_cand_report *report;          // optional [n_frames][REPORT_CH][NCAND]
    uint32_t *err;                    // device error flag
    uint32_t *counters;               // [0] loud frames, [1] pass-3 rounds, [2] LPC sizes from the window, [3] window misses,
                                      // [4] fixed candidates evaluated exactly, [5] candidates pruned by bounds
    uint32_t smem_plane_bytes;        // bytes of dynamic shared memory available for sample planes
    const uint32_t *crc_tab;          // CRC tables in global memory (crc_tables_host layout), staged into shared memory by every CTA
    uint32_t *frame_crc;              // out: raw CRC state R(frame bytes) of every frame (folded per track by k_crc_frames)
    uint32_t stagger;                 // SM clocks the second half of the CTAs waits before its first frame (de-phases the CTAs that share an SM)
    uint32_t work_bytes;              // bytes of the CTA's work area in shared memory (ingest stages / packer ring), multiple of 128, >= 16 KB
    uint32_t defer_bytes;             // per-CTA stride of defer_scratch; a frame of fsize + 64 <= defer_bytes is packed there first (0: never)
    uint8_t *defer_scratch;           // grid x defer_bytes: frames packed before their output offset is known (small frames, see k_encode_frames)
    unsigned long long *phase_cycles; // [0] ingest, [1] analysis, [2] look-back, [3] pack, [4] whole frame (SM clocks, thread 0)
};

struct FinalParams {
    const TrackDev *tracks;
    uint32_t n_tracks;
    uint32_t n_frames;
    int level;
    const uint2 *frames;
    uint8_t *out;
    const uint8_t *meta;              // metadata arena
    const unsigned long long *frame_excl;
    const uint32_t *frame_size;
    uint32_t *track_crc;              // per track, zeroed; the frames' shifted CRCs are XORed in
    const uint32_t *frame_crc;        // raw CRC state of every frame's bytes (written by the encode kernel)
    uint32_t n_segs;
    unsigned long long *file_off;     // out: per track
    unsigned long long *file_len;     // out: per track
};

// kernel launchers (flo_kernels.cu)
struct EncodeVariant {
    int threads;                      // threads per CTA
    int ctas_per_sm;                  // CTAs this variant is built to co-reside per SM (register budget)
    int ctas_fixed;                   // the same for the instantiation without LPC (levels 0-3), which needs fewer registers
    size_t (*static_smem)();          // bytes of the kernel's own shared state (in front of the planes)
    cudaError_t (*configure)(size_t dyn_smem);
    cudaError_t (*launch)(const EncodeParams &p, int grid, size_t dyn_smem, cudaStream_t st);
    int (*occupancy)(size_t dyn_smem);
};
const EncodeVariant &encode_variant(int threads);     // 512, 256 or 128
cudaError_t launch_setup(const TrackDev *tracks, uint32_t n_tracks, uint2 *frames, uint32_t n_frames, cudaStream_t st);
cudaError_t launch_toc(const FinalParams &p, cudaStream_t st);
cudaError_t launch_crc_frames(const FinalParams &p, cudaStream_t st);
// host copy of the CRC tables the encode kernel wants: slice0[256] | per NT in {128, 256, 512}: xop[4][256], klane[NT]
constexpr int CRC_TAB_WORDS = 256 + 3 * 1024 + 128 + 256 + 512;
__host__ __device__ constexpr int crc_tab_offset(int nt) { return 256 + (nt == 128 ? 0 : nt == 256 ? 1024 + 128 : 2 * 1024 + 128 + 256); }
void crc_tables_host(uint32_t *out);
cudaError_t launch_headers(const FinalParams &p, cudaStream_t st);
void upload_crc_tables();
// reflo's U8 / S32 ingest arms (reflo/src/audio.rs:255-269) as a pre-pass: interleaved PCM -> interleaved f32
cudaError_t launch_ingest_convert(const void *src, float *dst, unsigned long long n, int format, cudaStream_t st);

// waveform peaks of libflo::encode()'s analysis metadata (libflo/src/core/analysis.rs:38-119)
struct PeakParams {
    const float *x;                   // interleaved f32 samples
    unsigned long long n;             // samples.len()
    double spp;                       // samples_per_peak = sample_rate / peaks_per_second, in f64
    uint32_t channels;
    unsigned long long n_peaks;       // windows that start inside the input
    float *peaks;                     // out
    unsigned *max_bits;               // scratch: bit pattern of the largest peak
};
cudaError_t launch_waveform_peaks(const PeakParams &p, cudaStream_t st);     // 2 kernels

// K-weighted hop energies for the EBU R128 integrated loudness of libflo::encode()'s metadata (core/ebu_r128.rs)
struct KwParams {
    const float *x;                   // interleaved f32 samples
    unsigned long long frames;        // sample frames per channel
    uint32_t channels;
    uint32_t hop;                     // frames per 100 ms hop (= one segment of the filter)
    unsigned long long n_hops;        // ceil(frames / hop)
    double co[10];                    // shelf b0 b1 b2 a1 a2 | high-pass b0 b1 b2 a1 a2 (ebu_r128.rs:58-102)
    double *state;                    // [channels][n_hops][4]: pass 1 writes the segment's final state from a zero start,
                                      // pass 2 reads the true start state from the same slots (rewritten by the host scan)
    double *hop_sum;                  // [channels][n_hops]: sum of y^2 over the hop (pass 2)
};
cudaError_t launch_kweight(const KwParams &p, int pass, cudaStream_t st);

// ---- lossless decoder (libflo/src/reader.rs + libflo/src/lossless/decoder.rs) ----
enum DecErr : uint32_t { DEC_TOO_MANY = 1, DEC_BAD_ORDER = 2, DEC_EOF = 3, DEC_TRANSFORM = 4, DEC_BAD_K = 5 };
struct DecFrame { uint32_t type_flags; uint32_t n; };          // type | flags << 8 ; frame_samples (0 past the reader's break)
struct DecUnit  { unsigned long long pos; uint32_t size; uint32_t pad; };   // one channel of one frame: payload position and ch_size
struct DecodeParams {
    const uint8_t *file;              // file image in device memory; readable up to the next 16-byte boundary after file + len
    unsigned long long len;
    unsigned long long toc_pos;       // first TOC entry (host checked that the TOC lies inside the file)
    uint32_t n_toc;
    uint32_t channels;
    unsigned long long data_start, data_end;
    DecFrame *frames;                 // [n_toc]
    DecUnit *units;                   // [n_toc * channels]
    unsigned long long *base;         // [n_toc] exclusive prefix of frame_samples
    uint32_t *ctl;                    // [0] index of the first TOC entry the reader breaks on (init n_toc), [1] smallest error key
                                      // (init 0xFFFFFFFF; frame << 13 | (channel + 1) << 4 | DecErr), [2..3] total sample frames (u64)
    void *out;                        // interleaved samples, total * channels: f32, or i16 when out_i16
    int out_i16;                      // 1: the integer samples before i32_to_f32, saturated to i16 (flo_decode_i16)
};
cudaError_t launch_decode_parse(const DecodeParams &p, cudaStream_t st);   // parse + scan (2 kernels)
cudaError_t launch_decode_units(const DecodeParams &p, cudaStream_t st);   // 1 kernel


}  // namespace flo
