// flo_kernels.cu -- setup / TOC / CRC / header kernels and the variant table of the lossless ALPC encode path
// (the frame-encode kernel itself is encode_v3_body.cuh, built three times by flo_encode_nt*.cu).
//
// One persistent CTA per SM takes frames from a ticket counter (ticket order ==
// global frame order, which makes the decoupled look-back below deadlock free),
// keeps the frame's 16-bit channel planes resident in shared memory and runs the
// whole per-frame pipeline of the reference on them:
//
//   ingest      f32 (or i16 PCM) -> i32 quantise, silence test, deinterleave,
//               mid/side energies            encoder.rs:66-100, 131-170; audio_constants.rs:18-20
//   analyse     per channel: exact size of every candidate the reference tries
//               (raw, fixed 0..4, LPC 5..P)  encoder.rs:173-287; lpc.rs:213-359; rice.rs:29-69
//   select      strictly-smallest, first wins encoder.rs:184-216; frame typing encoder.rs:102-127
//   scan        frame offset = exclusive prefix of frame sizes (decoupled look-back)  writer.rs:199-220
//   pack        Rice / raw bit packing of the winner, MSB-first  rice.rs:84-114, 162-208; writer.rs:236-301
//
// TOC, CRC32 (per-segment CRC + GF(2) combine) and the 70-byte header are
// written by three small kernels afterwards (writer.rs:39-224, crc32.rs).
//
// Everything is integer-exact; the only floating point is the f32 quantiser
// (one RN multiply) and the sequential f64 Levinson-Durbin recursion, both with
// explicit _rn intrinsics so that no FMA contraction can change a bit.
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "flo_internal.h"

namespace flo {

typedef unsigned long long u64;
typedef long long i64;
typedef uint32_t u32;
typedef int32_t i32;

// ----------------------------------------------------------------------------
// constant tables (CRC)
// ----------------------------------------------------------------------------
__constant__ u32 c_crc_slice[4][256];   // slice-by-4 tables of the reflected CRC-32 (crc32.rs:2-20)
__constant__ u32 c_x2n[32];             // x^(2^i) mod p, reflected (for crc combine)
__constant__ u32 c_xop[4][256];         // v -> v * x^1024 mod p, by byte of v (strided Horner step of the CRC kernel)
__constant__ u32 c_klane[32];           // x^(32 (32 - lane)) mod p

static u32 h_multmodp(u32 a, u32 b) {
    u32 m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
static u32 h_x2n[32];
static u32 h_x2nmodp(u64 n, unsigned k) {
    u32 p = 1u << 31;
    while (n) { if (n & 1) p = h_multmodp(h_x2n[k & 31], p); n >>= 1; k++; }
    return p;
}

static u32 h_tab[CRC_TAB_WORDS];
void crc_tables_host(uint32_t *out) { upload_crc_tables(); memcpy(out, h_tab, sizeof h_tab); }

void upload_crc_tables() {
    static u32 slice[4][256];
    for (u32 i = 0; i < 256; i++) {
        u32 c = i;
        for (int j = 0; j < 8; j++) c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
        slice[0][i] = c;
    }
    for (u32 i = 0; i < 256; i++)
        for (int s = 1; s < 4; s++) slice[s][i] = (slice[s - 1][i] >> 8) ^ slice[0][slice[s - 1][i] & 0xFF];
    u32 p = 1u << 30;                      // x^1
    h_x2n[0] = p;
    for (int n = 1; n < 32; n++) h_x2n[n] = p = h_multmodp(p, p);
    static u32 xop[4][256];
    const u32 x1024 = h_x2nmodp(128, 3);
    for (int b = 0; b < 4; b++)
        for (u32 t = 0; t < 256; t++) xop[b][t] = h_multmodp(x1024, t << (8 * b));
    u32 klane[32];
    for (int l = 0; l < 32; l++) klane[l] = h_x2nmodp(4 * (32 - l), 3);
    // tables of the encode kernel's per-frame CRC (one stream per thread of the CTA, see encode_v3_body.cuh):
    // slice0[256] | for NT in 128, 256, 512: xop_NT[4][256] (v -> v * x^(32 NT)), klane_NT[NT] (x^(32 (NT - t)))
    memcpy(h_tab, slice[0], 1024);
    u32 *w = h_tab + 256;
    for (int nt = 128; nt <= 512; nt *= 2) {
        const u32 xs = h_x2nmodp((u64)4 * nt, 3);
        for (int b = 0; b < 4; b++)
            for (u32 t = 0; t < 256; t++) *w++ = h_multmodp(xs, t << (8 * b));
        for (int t = 0; t < nt; t++) *w++ = h_x2nmodp((u64)4 * (nt - t), 3);
    }
    cudaMemcpyToSymbol(c_crc_slice, slice, sizeof slice);
    cudaMemcpyToSymbol(c_x2n, h_x2n, sizeof h_x2n);
    cudaMemcpyToSymbol(c_xop, xop, sizeof xop);
    cudaMemcpyToSymbol(c_klane, klane, sizeof klane);
}

__device__ __forceinline__ u32 multmodp(u32 a, u32 b) {
    u32 m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}

// ----------------------------------------------------------------------------
// setup / finalise kernels
// ----------------------------------------------------------------------------
__global__ void k_setup_frames(const TrackDev *tracks, u32 n_tracks, uint2 *frames, u32 n_frames) {
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_frames) return;
    u32 lo = 0, hi = n_tracks;            // last track with first_frame <= g and n_frames > 0
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (tracks[mid].first_frame <= g) lo = mid; else hi = mid;
    }
    frames[g] = make_uint2(lo, g - tracks[lo].first_frame);
}

__device__ __forceinline__ void put_u32le(uint8_t *p, u32 v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

__device__ __forceinline__ u64 excl_at(const FinalParams &p, u32 g) {
    if (g < p.n_frames) return p.frame_excl[g];
    return p.n_frames ? p.frame_excl[p.n_frames - 1] + p.frame_size[p.n_frames - 1] : 0ull;
}

// build_toc_chunk, writer.rs:193-224: one thread per frame
__global__ void k_write_toc(const FinalParams p) {
    const u32 g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.n_frames) return;
    const uint2 fd = p.frames[g];
    const TrackDev tr = p.tracks[fd.x];
    const u32 f = fd.y;
    const u64 e0 = excl_at(p, tr.first_frame);
    const u64 file0 = tr.static_off + e0;
    uint8_t *e = p.out + file0 + FILE_HDR + 4 + 20ull * f;
    put_u32le(e, f);
    const u64 off = p.frame_excl[g] - e0;
    put_u32le(e + 4, (u32)off);
    put_u32le(e + 8, (u32)(off >> 32));
    put_u32le(e + 12, p.frame_size[g]);
    // cumulative samples before frame f = f * sample_rate (all earlier frames are full, encoder.rs:53-58)
    const u64 ts = ((u64)f * tr.sample_rate) * 1000ull / tr.sample_rate;
    put_u32le(e + 16, (u32)ts);
}

// CRC32 of the DATA chunk (crc32.rs:23-30, writer.rs:58).  The raw CRC R(M) = M(x) x^32 mod p is linear and
// ignores leading zeros, so the DATA chunk's CRC is the XOR of every frame's R shifted by the bytes that
// follow the frame in the chunk.  The encode kernel leaves R(frame) in frame_crc (computed from the frame's bytes
// while they are still in L2); this kernel only applies the shifts: one warp per frame.
// multmodp above and the x^(2^k) table follow zlib's crc32_combine helpers (zlib 1.2.12+, crc32.c; (C) 1995-2022 Mark Adler,
// zlib licence), re-typed here.
// x^(8 n) mod p for a byte count n, by a whole warp: lane b holds the factor of bit b of n (x^(8 * 2^b) from
// c_x2n, whose index wraps with period 32 like zlib's x2nmodp), five rounds of multmodp fold the 32 factors.
// The one-thread form walks up to 32 dependent multmodp's (40 us for the 3600 frames of an hour, divergent).
__device__ __forceinline__ u32 warp_x8n(u64 n) {
    const unsigned lane = threadIdx.x & 31u;
    const u32 t = c_x2n[(lane + 3u) & 31u];
    u32 f = ((n >> lane) & 1ull) ? t : 0x80000000u;                 // 0x80000000 = x^0
    if ((n >> (32u + lane)) & 1ull) f = multmodp(f, t);              // bit 32 + b: same table entry
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) f = multmodp(f, __shfl_xor_sync(0xffffffffu, f, o));
    return f;
}
// one warp per frame
__global__ void k_crc_frames(const FinalParams p) {
    const u32 g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= p.n_frames) return;
    const uint2 fd = p.frames[g];
    const TrackDev tr = p.tracks[fd.x];
    const u64 e1 = excl_at(p, tr.first_frame + tr.n_frames);
    const u64 after = e1 - (p.frame_excl[g] + p.frame_size[g]);
    const u32 sh = warp_x8n(after);
    if ((threadIdx.x & 31u) == 0) atomicXor(&p.track_crc[fd.x], multmodp(sh, p.frame_crc[g]));
}

// write_header_ex, writer.rs:132-191, + metadata (writer.rs:96): one block per track
__global__ void k_write_headers(const FinalParams p) {
    const u32 t = blockIdx.x;
    const TrackDev tr = p.tracks[t];
    const u64 e0 = excl_at(p, tr.first_frame), e1 = excl_at(p, tr.first_frame + tr.n_frames);
    const u64 dsize = e1 - e0;
    const u64 file0 = tr.static_off + e0;
    const u64 toc_size = 4 + 20ull * tr.n_frames;
    uint8_t *o = p.out + file0;
    u32 xd = 0;
    if (threadIdx.x < 32) xd = warp_x8n(dsize);
    if (threadIdx.x == 0) {
        // crc(M) = ~(R(M) ^ 0xFFFFFFFF x^(8 |M|)): initial state and final complement of crc32.rs:24-29
        const u32 crc = ~(p.track_crc[t] ^ multmodp(xd, 0xFFFFFFFFu));
        o[0] = 0x46; o[1] = 0x4C; o[2] = 0x4F; o[3] = 0x21;        // "FLO!", types.rs:6
        o[4] = 1; o[5] = 2;                                        // version 1.2, types.rs:12-13
        o[6] = 0; o[7] = 0;                                        // flags: lossless
        put_u32le(o + 8, tr.sample_rate);
        o[12] = (uint8_t)tr.channels;
        o[13] = (uint8_t)tr.bit_depth;
        const u64 total = tr.n_inter / tr.channels;                // sum of frame_samples
        put_u32le(o + 14, (u32)total); put_u32le(o + 18, (u32)(total >> 32));
        o[22] = (uint8_t)p.level; o[23] = 0; o[24] = 0; o[25] = 0;
        put_u32le(o + 26, crc);
        const u64 v[5] = {66ull, toc_size, dsize, 0ull, tr.meta_len};
        for (int i = 0; i < 5; i++) { put_u32le(o + 30 + 8 * i, (u32)v[i]); put_u32le(o + 34 + 8 * i, (u32)(v[i] >> 32)); }
        put_u32le(o + FILE_HDR, tr.n_frames);                      // TOC entry count, writer.rs:196
        p.file_off[t] = file0;
        p.file_len[t] = FILE_HDR + toc_size + dsize + tr.meta_len;
    }
    uint8_t *m = o + FILE_HDR + toc_size + dsize;
    for (u64 i = threadIdx.x; i < tr.meta_len; i += blockDim.x) m[i] = p.meta[tr.meta_off + i];
}

// ----------------------------------------------------------------------------
// launchers
// ----------------------------------------------------------------------------
extern const EncodeVariant g_variant_nt512, g_variant_nt256, g_variant_nt128;
const EncodeVariant &encode_variant(int threads) {
    return threads == 128 ? g_variant_nt128 : (threads == 256 ? g_variant_nt256 : g_variant_nt512);
}
cudaError_t launch_setup(const TrackDev *tracks, uint32_t n_tracks, uint2 *frames, uint32_t n_frames, cudaStream_t st) {
    if (n_frames == 0) return cudaSuccess;
    k_setup_frames<<<(n_frames + 255) / 256, 256, 0, st>>>(tracks, n_tracks, frames, n_frames);
    return cudaGetLastError();
}
cudaError_t launch_toc(const FinalParams &p, cudaStream_t st) {
    if (p.n_frames == 0) return cudaSuccess;
    k_write_toc<<<(p.n_frames + 255) / 256, 256, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_crc_frames(const FinalParams &p, cudaStream_t st) {
    if (p.n_frames == 0) return cudaSuccess;
    k_crc_frames<<<(p.n_frames + 7) / 8, 256, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_headers(const FinalParams &p, cudaStream_t st) {
    if (p.n_tracks == 0) return cudaSuccess;
    k_write_headers<<<p.n_tracks, 128, 0, st>>>(p);
    return cudaGetLastError();
}



// reflo's remaining ingest arms (append_samples, reflo/src/audio.rs:255-269), element-wise and exact in f32:
//   S32: s as f32 * (1.0 / 2147483648.0)        U8: (s as f32 - 128.0) / 128.0
// A separate pre-pass (one extra read + write of the samples) so that the encode kernel keeps its two ingest forms.
__global__ void k_ingest_convert(const void *src, float *dst, unsigned long long n, int format) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float v;
        if (format == FLO_FMT_S32) v = __fmul_rn(__int2float_rn(reinterpret_cast<const int32_t *>(src)[i]), 1.0f / 2147483648.0f);
        else v = __fdiv_rn(__fsub_rn((float)reinterpret_cast<const uint8_t *>(src)[i], 128.0f), 128.0f);
        dst[i] = v;
    }
}
cudaError_t launch_ingest_convert(const void *src, float *dst, unsigned long long n, int format, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    const unsigned long long want = (n + 255) / 256;
    const unsigned grid = (unsigned)(want < 148ull * 16 ? want : 148ull * 16);
    k_ingest_convert<<<grid, 256, 0, st>>>(src, dst, n, format);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// libflo::encode()'s waveform peaks (libflo/src/core/analysis.rs:38-119), exact in f32:
// window idx covers the sample frames [trunc(idx * spp), trunc((idx + 1) * spp)) with spp = sample_rate /
// peaks_per_second in f64 (:52, :58-62); mono takes max |s| (:73-78), stereo the mean of the two channels'
// max |s| over whole pairs (:80-91), any other channel count the maximum of the per-frame sample means, without
// abs (:93-99); all peaks are then divided by the largest one (:104-110).  max, add, divide are IEEE operations,
// so the result does not depend on the order the window is read in -- except the sequential f32 sum of the
// many-channel arm, which one lane does in the reference's order.  One warp per window; the largest peak is
// kept as the maximum of the bit patterns (peaks are >= 0).
// ------------------------------------------------------------------------------------------------
__global__ void k_waveform_peaks(const PeakParams p) {
    const unsigned lane = threadIdx.x & 31u;
    const u64 nwarps = ((u64)gridDim.x * blockDim.x) >> 5;
    for (u64 idx = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; idx < p.n_peaks; idx += nwarps) {
        u64 s = __double2ull_rz(__dmul_rn((double)idx, p.spp));
        u64 e = __double2ull_rz(__dmul_rn(__dadd_rn((double)idx, 1.0), p.spp));
        s *= p.channels;
        e = min(e * (u64)p.channels, p.n);
        const float *w = p.x + s;
        const u64 len = e > s ? e - s : 0;
        float peak;
        if (p.channels == 1) {
            float m = 0.0f;
#pragma unroll 4
            for (u64 i = lane; i < len; i += 32) m = fmaxf(m, fabsf(__ldg(w + i)));
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            peak = m;
        } else if (p.channels == 2) {
            float l = 0.0f, r = 0.0f;
            if (((uintptr_t)w & 7u) == 0) {                    // pairs as 8-byte loads (the window starts on a left sample)
                const float2 *w2 = reinterpret_cast<const float2 *>(w);
#pragma unroll 4
                for (u64 i = lane; i < (len >> 1); i += 32) { const float2 q = __ldg(w2 + i); l = fmaxf(l, fabsf(q.x)); r = fmaxf(r, fabsf(q.y)); }
            } else {
                for (u64 i = lane; i < (len >> 1); i += 32) { l = fmaxf(l, fabsf(w[2 * i])); r = fmaxf(r, fabsf(w[2 * i + 1])); }
            }
            for (int o = 16; o > 0; o >>= 1) { l = fmaxf(l, __shfl_xor_sync(0xffffffffu, l, o)); r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o)); }
            peak = __fdiv_rn(__fadd_rn(l, r), 2.0f);
        } else {
            float m = 0.0f;
            const u64 C = p.channels, nchunk = (len + C - 1) / C;
            for (u64 q = lane; q < nchunk; q += 32) {
                const u64 a = q * C, b = min(a + C, len);
                float sum = 0.0f;
                for (u64 i = a; i < b; i++) sum = __fadd_rn(sum, w[i]);
                m = fmaxf(m, __fdiv_rn(sum, (float)(b - a)));
            }
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            peak = m;
        }
        if (lane == 0) {
            p.peaks[idx] = peak;
            if (peak > 0.0f) atomicMax(p.max_bits, __float_as_uint(peak));
        }
    }
}
__global__ void k_normalise_peaks(float *peaks, u64 n, const unsigned *max_bits) {
    const float mx = __uint_as_float(*max_bits);
    if (!(mx > 0.0f)) return;                                                 // analysis.rs:105
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) peaks[i] = __fdiv_rn(peaks[i], mx);
}
cudaError_t launch_waveform_peaks(const PeakParams &p, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(p.max_bits, 0, 4, st);
    if (e != cudaSuccess || p.n_peaks == 0) return e;
    const u64 want = (p.n_peaks + 7) / 8;                                     // 8 warps (windows) per CTA
    k_waveform_peaks<<<(unsigned)(want < 148ull * 8 ? want : 148ull * 8), 256, 0, st>>>(p);
    const u64 w2 = (p.n_peaks + 255) / 256;
    k_normalise_peaks<<<(unsigned)(w2 < 148ull ? w2 : 148ull), 256, 0, st>>>(p.peaks, p.n_peaks, p.max_bits);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// EBU R128 integrated loudness of libflo::encode()'s metadata (libflo/src/core/ebu_r128.rs):
// the K-weighting (two cascaded biquads per channel, :43-48, :105-109) is a serial recurrence over the whole
// channel in the reference.  Here a channel is cut into its 100 ms hops; pass 1 runs every hop from a zero state
// and keeps the final state (the forced response), the host chains the hops' start states with the 4 x 4
// transition matrix of one hop (s' = M s + forced), and pass 2 runs every hop again from its true start state and
// sums y^2 -- exactly the reference's operations inside a hop (separately rounded products and sums, same
// order), so the only differences from the sequential result are the rounding of the chained start states and
// the regrouping of a block's sum into its four hops: ~1e-13 LU (tolerance stated in tests/test_gpu_analysis.py).
// One thread per (channel, hop), a CTA of 32 hops x C channels.  The samples of the CTA's hops are staged through
// shared memory 32 frames at a time -- every warp reads whole 128-byte pieces of a hop's interleaved frames and
// stores them transposed ([float index within the tile][hop], row stride 33: conflict-free both ways) -- because a
// serial filter per thread would otherwise touch 32 different lines with every load of a warp.
// ------------------------------------------------------------------------------------------------
constexpr int KW_HOPS = 32;           // hops per CTA (= lanes of a warp; warp = channel)
constexpr int KW_T = 32;              // frames per staged tile
constexpr int KW_MAXC = 8;            // more channels than this: the unstaged kernel

struct KwFilter {
    double z1s = 0.0, z2s = 0.0, z1h = 0.0, z2h = 0.0, acc = 0.0;
    template <int PASS>
    __device__ __forceinline__ void step(const double (&co)[10], double v) {
        const double y1 = __dadd_rn(__dmul_rn(co[0], v), z1s);                                 // Biquad::process, shelf
        z1s = __dadd_rn(__dsub_rn(__dmul_rn(co[1], v), __dmul_rn(co[3], y1)), z2s);
        z2s = __dsub_rn(__dmul_rn(co[2], v), __dmul_rn(co[4], y1));
        const double y = __dadd_rn(__dmul_rn(co[5], y1), z1h);                                 // high-pass
        z1h = __dadd_rn(__dsub_rn(__dmul_rn(co[6], y1), __dmul_rn(co[8], y)), z2h);
        z2h = __dsub_rn(__dmul_rn(co[7], y1), __dmul_rn(co[9], y));
        if (PASS == 2) acc = __dadd_rn(acc, __dmul_rn(y, y));                                  // ebu_r128.rs:248-250
    }
};

template <int PASS>
__global__ void k_kweight_staged(const KwParams p) {
    extern __shared__ float kw_tile[];                     // [KW_T * C][33]
    const u32 C = p.channels;
    const u32 lane = threadIdx.x & 31u, ch = threadIdx.x >> 5;          // blockDim = 32 * C
    const u64 j0 = (u64)blockIdx.x * KW_HOPS, j = j0 + lane;
    const bool live = j < p.n_hops;
    const u64 tseg = (u64)ch * p.n_hops + j;
    double co[10];
#pragma unroll
    for (int i = 0; i < 10; i++) co[i] = p.co[i];
    KwFilter f;
    if (PASS == 2 && live) { const double *st = p.state + (tseg << 2); f.z1s = st[0]; f.z2s = st[1]; f.z1h = st[2]; f.z2h = st[3]; }
    const u64 my0 = j * p.hop, my1 = live ? min(my0 + p.hop, p.frames) : my0;
    const u32 tid = threadIdx.x;                           // blockDim = 32 * C = floats of one hop's tile
    const bool cta_full = j0 + KW_HOPS <= p.n_hops && (j0 + KW_HOPS) * p.hop <= p.frames;
    const u64 hop_floats = (u64)p.hop * C;
    // stage: iteration h = hop h of the CTA, thread tid = float tid of its frames [base, base + KW_T); all 32 loads of
    // a thread are issued back to back, and the loads of tile b + 1 are in flight while tile b is filtered
    float v[KW_HOPS];
    auto load_tile = [&](u32 base) {
        const float *src = p.x + (j0 * p.hop + base) * C + tid;
        if (cta_full && base + KW_T <= p.hop) {
#pragma unroll
            for (int h = 0; h < KW_HOPS; h++) v[h] = __ldg(src + h * hop_floats);
        } else {
#pragma unroll
            for (int h = 0; h < KW_HOPS; h++) {
                const u64 hj = j0 + h;
                const u64 f0 = hj * p.hop + base;
                const u64 f1 = hj < p.n_hops ? min(min(f0 + KW_T, (hj + 1) * p.hop), p.frames) : f0;
                const u64 nfl = f1 > f0 ? (f1 - f0) * C : 0;
                v[h] = tid < nfl ? __ldg(src + h * hop_floats) : 0.0f;
            }
        }
    };
    load_tile(0);
    for (u32 base = 0; base < p.hop; base += KW_T) {
        __syncthreads();
#pragma unroll
        for (int h = 0; h < KW_HOPS; h++) kw_tile[tid * 33 + h] = v[h];
        __syncthreads();
        if (base + KW_T < p.hop) load_tile(base + KW_T);
        const u64 fb = my0 + base;
        const u32 nst = fb < my1 ? (u32)min((u64)KW_T, my1 - fb) : 0u;
        if (nst == KW_T) {
#pragma unroll 8
            for (u32 i = 0; i < KW_T; i++) f.step<PASS>(co, (double)kw_tile[(i * C + ch) * 33 + lane]);
        } else {
            for (u32 i = 0; i < nst; i++) f.step<PASS>(co, (double)kw_tile[(i * C + ch) * 33 + lane]);
        }
    }
    if (!live) return;
    if (PASS == 1) { double *st = p.state + (tseg << 2); st[0] = f.z1s; st[1] = f.z2s; st[2] = f.z1h; st[3] = f.z2h; }
    else p.hop_sum[tseg] = f.acc;
}

// unstaged form for more than KW_MAXC channels: one thread per (channel, hop), loads straight from global
template <int PASS>
__global__ void k_kweight(const KwParams p) {
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.n_hops * p.channels) return;
    const u32 ch = (u32)(t / p.n_hops);
    const u64 j = t % p.n_hops;
    const u64 i0 = j * p.hop, i1 = min(i0 + p.hop, p.frames);
    double *st = p.state + (t << 2);
    double co[10];
#pragma unroll
    for (int i = 0; i < 10; i++) co[i] = p.co[i];
    KwFilter f;
    if (PASS == 2) { f.z1s = st[0]; f.z2s = st[1]; f.z1h = st[2]; f.z2h = st[3]; }
    const float *x = p.x + i0 * p.channels + ch;
    for (u64 i = i0; i < i1; i++, x += p.channels) f.step<PASS>(co, (double)__ldg(x));
    if (PASS == 1) { st[0] = f.z1s; st[1] = f.z2s; st[2] = f.z1h; st[3] = f.z2h; }
    else p.hop_sum[t] = f.acc;
}
cudaError_t launch_kweight(const KwParams &p, int pass, cudaStream_t st) {
    const u64 n = p.n_hops * p.channels;
    if (n == 0) return cudaSuccess;
    if (p.channels <= KW_MAXC) {
        const unsigned grid = (unsigned)((p.n_hops + KW_HOPS - 1) / KW_HOPS), threads = 32u * p.channels;
        const size_t smem = sizeof(float) * KW_T * p.channels * 33;
        if (pass == 1) k_kweight_staged<1><<<grid, threads, smem, st>>>(p);
        else k_kweight_staged<2><<<grid, threads, smem, st>>>(p);
    } else {
        const unsigned grid = (unsigned)((n + 127) / 128);
        if (pass == 1) k_kweight<1><<<grid, 128, 0, st>>>(p);
        else k_kweight<2><<<grid, 128, 0, st>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace flo
