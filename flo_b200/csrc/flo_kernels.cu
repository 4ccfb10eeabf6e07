// flo_kernels.cu -- sm_100a kernels of the lossless ALPC encode path.
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
__constant__ u32 c_pow128[33];          // x^(8*128*j) mod p, j = 0..32
__constant__ u32 c_powseg;              // x^(8*CRC_SEG) mod p

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
    u32 pow128[33];
    for (int j = 0; j <= 32; j++) pow128[j] = h_x2nmodp((u64)128 * j, 3);
    u32 powseg = h_x2nmodp(CRC_SEG, 3);
    cudaMemcpyToSymbol(c_crc_slice, slice, sizeof slice);
    cudaMemcpyToSymbol(c_x2n, h_x2n, sizeof h_x2n);
    cudaMemcpyToSymbol(c_pow128, pow128, sizeof pow128);
    cudaMemcpyToSymbol(c_powseg, &powseg, sizeof powseg);
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
__device__ u32 x2nmodp(u64 n, unsigned k) {
    u32 p = 1u << 31;
    while (n) { if (n & 1) p = multmodp(c_x2n[k & 31], p); n >>= 1; k++; }
    return p;
}

// ----------------------------------------------------------------------------
// shared state of the frame-encode CTA
// ----------------------------------------------------------------------------
struct Smem {
    u64 red64[NWARP][16];
    u32 red32[NWARP][16];
    u64 tot64[16];
    u32 tot32[16];
    i64 ac[MAXORD + 1];
    i32 qc[MAXORD + 1][MAXORD];       // quantised LPC coefficients by order (lpc.rs:263-273)
    i32 lpc_ok[MAXORD + 1];
    i32 lpc_shift[MAXORD + 1];
    i32 cand_k[NCAND];
    i64 cand_size[NCAND];
    u64 fix_sum[5];
    u32 fix_max[5];
    i32 wcoef[MAXORD];                // winner's coefficients while packing
    u32 scan_warp[NWARP];
    u32 scan_total;
    u32 g;                            // current global frame
    i32 ms;                           // mid/side chosen (encoder.rs:94-100)
    u64 frame_excl;                   // exclusive prefix of frame sizes
    u32 ring[RING_WORDS];             // bit-packer staging ring (big-endian bit order words)
};

size_t encode_static_smem() { return sizeof(Smem); }

__device__ __forceinline__ u64 warp_sum64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ u32 warp_max32(u32 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sums / maxima into s.tot64 / s.tot32 (valid for all threads on return).
template <int N>
__device__ __forceinline__ void block_sum64(Smem &s, const u64 (&v)[N]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; i++) {
        u64 t = warp_sum64(v[i]);
        if (lane == 0) s.red64[w][i] = t;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        u64 t = 0;
        for (int j = 0; j < NWARP; j++) t += s.red64[j][threadIdx.x];
        s.tot64[threadIdx.x] = t;
    }
    __syncthreads();
}
template <int N>
__device__ __forceinline__ void block_max32(Smem &s, const u32 (&v)[N]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; i++) {
        u32 t = warp_max32(v[i]);
        if (lane == 0) s.red32[w][i] = t;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        u32 t = 0;
        for (int j = 0; j < NWARP; j++) t = max(t, s.red32[j][threadIdx.x]);
        s.tot32[threadIdx.x] = t;
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------
// scalar pieces of the reference
// ----------------------------------------------------------------------------
// f32_to_i32, core/audio_constants.rs:18-20: (x * 32767.0).clamp(-32768, 32767) as i32.
// cvt.rzi saturates and maps NaN to 0 exactly like Rust's `as i32`; clamping the
// truncated integer equals truncating the clamped float because both bounds are integers.
__device__ __forceinline__ i32 f32_to_i32(float x) {
    float y = __fmul_rn(x, 32767.0f);
    int v = __float2int_rz(y);
    return max(-32768, min(32767, v));
}
// silence test of encoder.rs:70: |x| < 1e-7 (NaN is not silent)
__device__ __forceinline__ bool is_loud(float x) { return !(fabsf(x) < 1e-7f); }

template <typename T> __device__ __forceinline__ float sample_f32(const T *p, size_t i);
template <> __device__ __forceinline__ float sample_f32<float>(const float *p, size_t i) { return __ldg(p + i); }
// reflo/src/audio.rs:247-254: s as f32 * (1.0 / 32768.0)
template <> __device__ __forceinline__ float sample_f32<int16_t>(const int16_t *p, size_t i) {
    return __fmul_rn((float)__ldg(p + i), 1.0f / 32768.0f);
}

// estimate_rice_parameter_i32, core/rice.rs:29-69
__device__ __forceinline__ int rice_k(u32 max_abs, u64 sum_abs, u32 n) {
    if (n == 0) return 4;
    if (max_abs == 0) return 0;
    u64 mu = 2ull * max_abs;
    int min_k = 0;
    if (mu > 255) { int bits = 64 - __clzll((i64)mu); min_k = bits > 8 ? bits - 8 : 0; }
    u32 mean = (u32)(sum_abs / (u64)n);
    int mean_k = mean > 0 ? 32 - __clz((int)mean) : 0;
    int k = max(min_k, mean_k);
    return min(k, 15);
}
__device__ __forceinline__ u32 zigzag(i32 r) { return ((u32)r << 1) ^ (u32)(r >> 31); }   // rice.rs:96
__device__ __forceinline__ u32 uabs(i32 r) { return r < 0 ? 0u - (u32)r : (u32)r; }

// lpc_order_from_level, encoder.rs:289-302
__device__ __forceinline__ int order_of_level(int level) {
    const int t[10] = {0, 2, 4, 4, 6, 8, 8, 10, 12, 12};
    return t[level < 0 ? 0 : (level > 9 ? 9 : level)];
}

// levinson_durbin_int, lpc.rs:225-276 -- sequential f64, every product and sum
// rounded separately.  The recursion is prefix consistent (the order-m result is
// the state after iteration m-1), so one run to order P yields every order 5..P.
__device__ void levinson_all_orders(Smem &s, int P) {
    for (int o = 0; o <= MAXORD; o++) { s.lpc_ok[o] = 0; s.lpc_shift[o] = 0; }
    if (s.ac[0] == 0) return;
    double a[MAXORD], nc[MAXORD];
    for (int i = 0; i < MAXORD; i++) a[i] = 0.0;
    double err = (double)s.ac[0];
    for (int i = 0; i < P; i++) {
        double lambda = (double)s.ac[i + 1];
        for (int j = 0; j < i; j++) lambda = __dsub_rn(lambda, __dmul_rn(a[j], (double)s.ac[i - j]));
        if (fabs(err) < 1e-10) return;
        double gamma = __ddiv_rn(lambda, err);
        if (fabs(gamma) >= 1.0) return;
        nc[i] = gamma;
        for (int j = 0; j < i; j++) nc[j] = __dsub_rn(a[j], __dmul_rn(gamma, a[i - 1 - j]));
        for (int j = 0; j <= i; j++) a[j] = nc[j];
        err = __dmul_rn(err, __dsub_rn(1.0, __dmul_rn(gamma, gamma)));
        const int o = i + 1;
        if (o >= 5) {
            double mx = 0.0;
            for (int j = 0; j < o; j++) { double t = fabs(a[j]); if (t == t && t > mx) mx = t; }
            if (mx == 0.0 || isinf(mx)) continue;
            // shift = min(floor(log2(2^30 / max)) as u8, 15); floor(log2(v)) of a positive finite
            // double is its binary exponent (|a_j| <= C(12,6) = 924 makes this 15 in practice).
            double v = __ddiv_rn(1073741824.0, mx);
            int e = isinf(v) ? 255 : ilogb(v);
            int shift = e < 0 ? 0 : (e > 15 ? 15 : e);
            double scale = (double)(1ll << shift);
            for (int j = 0; j < o; j++) {
                double q = round(__dmul_rn(a[j], scale));       // f64::round: half away from zero
                i32 qi = q >= 2147483647.0 ? 2147483647 : (q <= -2147483648.0 ? (-2147483647 - 1) : (i32)q);
                s.qc[o][j] = qi;
            }
            s.lpc_shift[o] = shift;
            s.lpc_ok[o] = 1;
        }
    }
}

// ----------------------------------------------------------------------------
// sample access: 16 samples of the coded channel starting at i0 (multiple of 16)
// plus NH samples of history (zero before the frame start; planes are zero padded
// behind the channel end).  x[NH + j] = s[i0 + j], x[NH - 1 - h] = s[i0 - 1 - h].
// ----------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const int4 v, i32 *t) {
    t[0] = (i32)(int16_t)(v.x & 0xffff); t[1] = v.x >> 16;
    t[2] = (i32)(int16_t)(v.y & 0xffff); t[3] = v.y >> 16;
    t[4] = (i32)(int16_t)(v.z & 0xffff); t[5] = v.z >> 16;
    t[6] = (i32)(int16_t)(v.w & 0xffff); t[7] = v.w >> 16;
}
template <int NH>
__device__ __forceinline__ void load_plane(const int16_t *pl, int i0, i32 (&x)[NH + CH]) {
    i32 t[32];
    const int4 *p = reinterpret_cast<const int4 *>(pl + i0);
    const int4 z = make_int4(0, 0, 0, 0);
    unpack8(p[0], t + 16);
    unpack8(p[1], t + 24);
    if (NH > 8) unpack8(i0 > 0 ? p[-2] : z, t);
    if (NH > 0) unpack8(i0 > 0 ? p[-1] : z, t + 8);
#pragma unroll
    for (int i = 0; i < NH + CH; i++) x[i] = t[16 - NH + i];
}
// msmode: 0 = plane pa as is; 1 = mid = L + R; 2 = side = L - R (encoder.rs:156-170, no shift)
template <int NH>
__device__ __forceinline__ void load_x(const int16_t *pa, const int16_t *pb, int msmode, int i0, i32 (&x)[NH + CH]) {
    load_plane<NH>(pa, i0, x);
    if (msmode) {
        i32 y[NH + CH];
        load_plane<NH>(pb, i0, y);
        if (msmode == 1) {
#pragma unroll
            for (int i = 0; i < NH + CH; i++) x[i] = x[i] + y[i];
        } else {
#pragma unroll
            for (int i = 0; i < NH + CH; i++) x[i] = x[i] - y[i];
        }
    }
}

// compile-time loop over LPC orders
template <int O, int P> struct ForOrders {
    template <class F> static __device__ __forceinline__ void run(F &&f) {
        f(std::integral_constant<int, O>{});
        if constexpr (O < P) ForOrders<O + 1, P>::run(f);
    }
};

// fixed_predictor_residuals, lpc.rs:301-359: r_o[i] = o-th difference for i >= o and the
// i-th difference for i < o.  Streams through one chunk; fn(j, r0..r4) per sample.
template <class F>
__device__ __forceinline__ void fixed_chunk(const i32 (&x)[4 + CH], bool first, F &&fn) {
    // difference state just before the chunk (zero history at the frame start)
    i32 xp = x[3];
    i32 p1 = x[3] - x[2];
    i32 p1b = x[2] - x[1], p1c = x[1] - x[0];
    i32 p2 = p1 - p1b, p2b = p1b - p1c;
    i32 p3 = p2 - p2b;
#pragma unroll
    for (int j = 0; j < CH; j++) {
        i32 d0 = x[4 + j];
        i32 d1 = d0 - xp;
        i32 d2 = d1 - p1;
        i32 d3 = d2 - p2;
        i32 d4 = d3 - p3;
        xp = d0; p1 = d1; p2 = d2; p3 = d3;
        i32 r2 = d2, r3 = d3, r4 = d4;
        if (j < 4 && first) {              // warm-up of lpc.rs:311-352
            if (j == 0) { d1 = d0; r2 = d0; r3 = d0; r4 = d0; }
            if (j == 1) { r2 = d1; r3 = d1; r4 = d1; }
            if (j == 2) { r3 = d2; r4 = d2; }
            if (j == 3) { r4 = d3; }
        }
        fn(j, d0, d1, r2, r3, r4);
    }
}

// calc_residuals_int, lpc.rs:279-298, for one chunk; x has 12 samples of history.
template <int O, class F>
__device__ __forceinline__ void lpc_chunk(const i32 (&x)[MAXORD + CH], bool first, const i32 *qc, int shift, F &&fn) {
    i32 q[O];
#pragma unroll
    for (int j = 0; j < O; j++) q[j] = qc[j];
#pragma unroll
    for (int j = 0; j < CH; j++) {
        i64 pred = 0;
#pragma unroll
        for (int t = 0; t < O; t++) pred += (i64)q[t] * (i64)x[MAXORD + j - 1 - t];
        pred >>= shift;
        i32 r = (i32)((u32)x[MAXORD + j] - (u32)(i32)pred);
        if (first && j < O) r = x[MAXORD + j];       // warm-up, lpc.rs:283-285
        fn(j, r);
    }
}

// ----------------------------------------------------------------------------
// per-channel analysis: exact k and encoded size of every candidate
// ----------------------------------------------------------------------------
struct ChanView {
    const int16_t *pa, *pb;
    int msmode;
    int n;
};

template <int P>     // P = LPC max order to analyse (0 = fixed predictors only)
__device__ void analyse_channel(Smem &s, const ChanView cv, int level, int fmax, flo_cand_report *rep) {
    const int tid = threadIdx.x;
    const int n = cv.n;
    const int nchunks = (n + CH - 1) / CH;

    // ---- pass A: fixed-predictor statistics + autocorrelation (lpc.rs:213-221) ----
    {
        u64 fsum[5] = {0, 0, 0, 0, 0};
        u32 fmx[5] = {0, 0, 0, 0, 0};
        u64 acc[P + 1];
#pragma unroll
        for (int l = 0; l <= P; l++) acc[l] = 0;
        for (int c = tid; c < nchunks; c += NT) {
            const int i0 = c * CH;
            const int nv = min(CH, n - i0);
            constexpr int NH = P > 4 ? P : 4;
            i32 x[NH + CH];
            load_x<NH>(cv.pa, cv.pb, cv.msmode, i0, x);
            if constexpr (P > 0) {
#pragma unroll
                for (int j = 0; j < CH; j++) {
#pragma unroll
                    for (int l = 0; l <= P; l++) acc[l] += (u64)((i64)x[NH + j] * (i64)x[NH + j - l]);
                }
            }
            i32 xf[4 + CH];
#pragma unroll
            for (int i = 0; i < 4 + CH; i++) xf[i] = x[NH - 4 + i];
            u32 cs[5] = {0, 0, 0, 0, 0};
            fixed_chunk(xf, i0 == 0, [&](int j, i32 r0, i32 r1, i32 r2, i32 r3, i32 r4) {
                if (j < nv) {
                    u32 a0 = uabs(r0), a1 = uabs(r1), a2 = uabs(r2), a3 = uabs(r3), a4 = uabs(r4);
                    cs[0] += a0; cs[1] += a1; cs[2] += a2; cs[3] += a3; cs[4] += a4;
                    fmx[0] = max(fmx[0], a0); fmx[1] = max(fmx[1], a1); fmx[2] = max(fmx[2], a2);
                    fmx[3] = max(fmx[3], a3); fmx[4] = max(fmx[4], a4);
                }
            });
#pragma unroll
            for (int o = 0; o < 5; o++) fsum[o] += cs[o];
        }
        block_sum64<5>(s, fsum);
        if (tid < 5) s.fix_sum[tid] = s.tot64[tid];
        block_max32<5>(s, fmx);
        if (tid < 5) s.fix_max[tid] = s.tot32[tid];
        if constexpr (P > 0) {
            block_sum64<P + 1>(s, acc);
            if (tid <= P) s.ac[tid] = (i64)s.tot64[tid];
        }
        __syncthreads();
        if (tid == 0) {
            for (int j = 0; j < NCAND; j++) { s.cand_k[j] = 0; s.cand_size[j] = -1; }
            s.cand_size[0] = 2ll * n;                                  // encode_raw, encoder.rs:220-226
            for (int o = 0; o <= 4; o++) s.cand_k[1 + o] = rice_k(s.fix_max[o], s.fix_sum[o], (u32)n);
            if constexpr (P > 0) levinson_all_orders(s, P);
        }
        __syncthreads();
    }

    // ---- pass B: exact Rice size of the fixed candidates: sum(u >> k) + n(1 + k) bits (rice.rs:97-113).
    // The 255 cap of rice.rs:103 never binds: k >= bitlen(2 max|r|) - 8 makes u >> k <= 255. ----
    {
        const int k0 = s.cand_k[1], k1 = s.cand_k[2], k2 = s.cand_k[3], k3 = s.cand_k[4], k4 = s.cand_k[5];
        u64 q[5] = {0, 0, 0, 0, 0};
        for (int c = tid; c < nchunks; c += NT) {
            const int i0 = c * CH;
            const int nv = min(CH, n - i0);
            i32 xf[4 + CH];
            load_x<4>(cv.pa, cv.pb, cv.msmode, i0, xf);
            u32 cs[5] = {0, 0, 0, 0, 0};
            fixed_chunk(xf, i0 == 0, [&](int j, i32 r0, i32 r1, i32 r2, i32 r3, i32 r4) {
                if (j < nv) {
                    cs[0] += zigzag(r0) >> k0; cs[1] += zigzag(r1) >> k1; cs[2] += zigzag(r2) >> k2;
                    cs[3] += zigzag(r3) >> k3; cs[4] += zigzag(r4) >> k4;
                }
            });
#pragma unroll
            for (int o = 0; o < 5; o++) q[o] += cs[o];
        }
        block_sum64<5>(s, q);
        if (tid == 0) {
            for (int o = 0; o <= fmax; o++) {
                u64 bits = s.tot64[o] + (u64)n * (u64)(1 + s.cand_k[1 + o]);
                s.cand_size[1 + o] = (i64)((bits + 7) >> 3);
            }
        }
        __syncthreads();
    }

    // ---- LPC candidates 5..P (encoder.rs:204-214, 254-287) ----
    if constexpr (P > 0) {
        // pass A: max|r| and sum|r| per order
        u64 lsum[P - 4];
        u32 lmx[P - 4];
#pragma unroll
        for (int i = 0; i < P - 4; i++) { lsum[i] = 0; lmx[i] = 0; }
        for (int c = tid; c < nchunks; c += NT) {
            const int i0 = c * CH;
            const int nv = min(CH, n - i0);
            i32 x[MAXORD + CH];
            load_x<MAXORD>(cv.pa, cv.pb, cv.msmode, i0, x);
            ForOrders<5, P>::run([&](auto oc) {
                constexpr int O = decltype(oc)::value;
                if (s.lpc_ok[O] && n > O) {
                    u32 cs = 0, mx = lmx[O - 5];
                    lpc_chunk<O>(x, i0 == 0, s.qc[O], s.lpc_shift[O], [&](int j, i32 r) {
                        if (j < nv) { u32 a = uabs(r); cs += a; mx = max(mx, a); }
                    });
                    lsum[O - 5] += cs;
                    lmx[O - 5] = mx;
                }
            });
        }
        block_sum64<P - 4>(s, lsum);
        block_max32<P - 4>(s, lmx);
        if (tid == 0) {
            for (int o = 5; o <= P; o++) {
                // n <= order, Levinson failure, or max|r| > 1_000_000 (encoder.rs:255-257, 262-272)
                if (!(s.lpc_ok[o] && n > o) || s.tot32[o - 5] > 1000000u) { s.lpc_ok[o] = 0; continue; }
                s.cand_k[1 + o] = rice_k(s.tot32[o - 5], s.tot64[o - 5], (u32)n);
            }
        }
        __syncthreads();
        // pass B: exact sizes
        u64 q[P - 4];
#pragma unroll
        for (int i = 0; i < P - 4; i++) q[i] = 0;
        for (int c = tid; c < nchunks; c += NT) {
            const int i0 = c * CH;
            const int nv = min(CH, n - i0);
            i32 x[MAXORD + CH];
            load_x<MAXORD>(cv.pa, cv.pb, cv.msmode, i0, x);
            ForOrders<5, P>::run([&](auto oc) {
                constexpr int O = decltype(oc)::value;
                if (s.lpc_ok[O]) {
                    const int k = s.cand_k[1 + O];
                    u32 cs = 0;
                    lpc_chunk<O>(x, i0 == 0, s.qc[O], s.lpc_shift[O], [&](int j, i32 r) {
                        if (j < nv) cs += zigzag(r) >> k;
                    });
                    q[O - 5] += cs;
                }
            });
        }
        block_sum64<P - 4>(s, q);
        if (tid == 0) {
            for (int o = 5; o <= P; o++) {
                if (!s.lpc_ok[o]) continue;
                u64 bits = s.tot64[o - 5] + (u64)n * (u64)(1 + s.cand_k[1 + o]);
                s.cand_size[1 + o] = (i64)((bits + 7) >> 3);
            }
        }
        __syncthreads();
    }
    (void)level;
    if (rep && tid < NCAND) { rep[tid].k = s.cand_k[tid]; rep[tid].pad = 0; rep[tid].size = s.cand_size[tid]; }
}

// ----------------------------------------------------------------------------
// bit packer
// ----------------------------------------------------------------------------
// Codes of one chunk for the winner.  MODE 0..4 fixed, 5..12 LPC, 13 raw.
template <int MODE>
__device__ __forceinline__ void chunk_codes(const ChanView cv, int i0, const i32 *qc, int shift, u32 (&u)[CH]) {
    if constexpr (MODE == 13) {
        i32 x[CH];
        load_x<0>(cv.pa, cv.pb, cv.msmode, i0, x);
#pragma unroll
        for (int j = 0; j < CH; j++) {            // (s as i16).to_le_bytes(), encoder.rs:222-224
            u32 v = (u32)x[j] & 0xffffu;
            u[j] = ((v & 0xff) << 8) | (v >> 8);
        }
    } else if constexpr (MODE <= 4) {
        i32 xf[4 + CH];
        load_x<4>(cv.pa, cv.pb, cv.msmode, i0, xf);
        fixed_chunk(xf, i0 == 0, [&](int j, i32 r0, i32 r1, i32 r2, i32 r3, i32 r4) {
            i32 r = MODE == 0 ? r0 : MODE == 1 ? r1 : MODE == 2 ? r2 : MODE == 3 ? r3 : r4;
            u[j] = zigzag(r);
        });
    } else {
        i32 x[MAXORD + CH];
        load_x<MAXORD>(cv.pa, cv.pb, cv.msmode, i0, x);
        lpc_chunk<MODE>(x, i0 == 0, qc, shift, [&](int j, i32 r) { u[j] = zigzag(r); });
    }
}

__device__ __forceinline__ void emit_chunk(u32 *ring, const u32 (&u)[CH], int nv, int k, bool raw, u64 start,
                                           u32 wlo, u32 whi) {
    u32 w = (u32)(start >> 5);
    int nb = (int)(start & 31);
    u64 acc = 0;
    bool firstw = true;
    auto out = [&](u32 word) {
        if (w >= wlo && w < whi) {
            if (firstw) atomicOr(&ring[w & (RING_WORDS - 1)], word);
            else ring[w & (RING_WORDS - 1)] = word;
        }
        firstw = false;
        w++;
    };
    auto put = [&](u32 v, int len) {
        acc = (acc << len) | v;
        nb += len;
        if (nb >= 32) { out((u32)(acc >> (nb - 32))); nb -= 32; }
    };
#pragma unroll
    for (int j = 0; j < CH; j++) {
        if (j < nv) {
            if (raw) {
                put(u[j], 16);
            } else {                               // encode_sample, rice.rs:94-114
                u32 q = u[j] >> k;
                u32 rem = u[j] & ((1u << k) - 1u);
                if (q <= 16) {
                    put((((1u << q) - 1u) << (k + 1)) | rem, (int)q + k + 1);
                } else {
                    while (q > 0) { u32 t = q < 24 ? q : 24; put((1u << t) - 1u, (int)t); q -= t; }
                    put(rem, k + 1);
                }
            }
        }
    }
    if (nb > 0) {
        u32 word = (u32)(acc << (32 - nb));
        if (w >= wlo && w < whi) atomicOr(&ring[w & (RING_WORDS - 1)], word);
    }
}

// Copy completed ring words [wa, wb) to the output and clear them.  Word w of the ring maps
// to the 4-byte aligned address abase + 4 w; only bytes inside [lo, hi) belong to this payload.
__device__ __forceinline__ void flush_ring(u32 *ring, uint8_t *obase, u64 abase, u64 lo, u64 hi, u32 wa, u32 wb) {
    for (u32 w = wa + threadIdx.x; w < wb; w += NT) {
        const u32 slot = w & (RING_WORDS - 1);
        const u32 v = ring[slot];
        ring[slot] = 0;
        const u64 a = abase + 4ull * w;
        if (a >= lo && a + 4 <= hi) {
            *reinterpret_cast<u32 *>(obase + a) = __byte_perm(v, 0, 0x0123);
        } else {
#pragma unroll
            for (int b = 0; b < 4; b++)
                if (a + b >= lo && a + b < hi) obase[a + b] = (uint8_t)(v >> (24 - 8 * b));
        }
    }
}

// Pack one channel's residual payload at byte offset `pos` (relative to obase, which is 4-byte
// aligned) -- encode_i32 / BitWriter (rice.rs:84-92, 162-208) or encode_raw (encoder.rs:220-226).
__device__ void pack_channel(Smem &s, const ChanView cv, const ChanResult &cr, int shift, uint8_t *obase, u64 pos,
                             u32 *err) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = cv.n;
    const int mode = cr.kind == 0 ? 13 : cr.order;
    const bool raw = cr.kind == 0;
    const int k = cr.k;
    const u64 abase = pos & ~3ull;
    const u64 lo = pos, hi = pos + cr.nbytes;
    u64 bitpos = (pos & 3ull) * 8ull;
    u32 wfl = 0;
    const int per_sc = NT * CH;
    const int nsc = (n + per_sc - 1) / per_sc;
    for (int sc = 0; sc < nsc; sc++) {
        const int i0 = sc * per_sc + tid * CH;
        const int nv = max(0, min(CH, n - i0));
        u32 u[CH];
        u32 tb = 0;
        if (nv > 0) {
            switch (mode) {
                case 0: chunk_codes<0>(cv, i0, s.wcoef, shift, u); break;
                case 1: chunk_codes<1>(cv, i0, s.wcoef, shift, u); break;
                case 2: chunk_codes<2>(cv, i0, s.wcoef, shift, u); break;
                case 3: chunk_codes<3>(cv, i0, s.wcoef, shift, u); break;
                case 4: chunk_codes<4>(cv, i0, s.wcoef, shift, u); break;
                case 5: chunk_codes<5>(cv, i0, s.wcoef, shift, u); break;
                case 6: chunk_codes<6>(cv, i0, s.wcoef, shift, u); break;
                case 7: chunk_codes<7>(cv, i0, s.wcoef, shift, u); break;
                case 8: chunk_codes<8>(cv, i0, s.wcoef, shift, u); break;
                case 9: chunk_codes<9>(cv, i0, s.wcoef, shift, u); break;
                case 10: chunk_codes<10>(cv, i0, s.wcoef, shift, u); break;
                case 11: chunk_codes<11>(cv, i0, s.wcoef, shift, u); break;
                case 12: chunk_codes<12>(cv, i0, s.wcoef, shift, u); break;
                default: chunk_codes<13>(cv, i0, s.wcoef, shift, u); break;
            }
            if (raw) {
                tb = 16u * (u32)nv;
            } else {
#pragma unroll
                for (int j = 0; j < CH; j++)
                    if (j < nv) tb += (u[j] >> k) + 1u + (u32)k;
            }
        }
        // block exclusive scan of the chunk bit counts
        u32 inc = tb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s.scan_warp[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            u32 v = lane < NWARP ? s.scan_warp[lane] : 0;
            u32 vi = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                u32 t = __shfl_up_sync(0xffffffffu, vi, o);
                if (lane >= o) vi += t;
            }
            if (lane < NWARP) s.scan_warp[lane] = vi - v;
            if (lane == NWARP - 1) s.scan_total = vi;
        }
        __syncthreads();
        const u64 start = bitpos + s.scan_warp[wid] + (inc - tb);
        const u64 end_sc = bitpos + s.scan_total;
        const u32 wlast = (u32)((end_sc + 31) >> 5);
        u32 wlo = wfl;
        for (;;) {
            const u32 whi = wlo + RING_WORDS;
            if (nv > 0) emit_chunk(s.ring, u, nv, k, raw, start, wlo, whi);
            __syncthreads();
            const u32 wend = min(whi, (u32)(end_sc >> 5));
            flush_ring(s.ring, obase, abase, lo, hi, wlo, wend);
            __syncthreads();
            wlo = wend;
            if (whi >= wlast) break;
        }
        wfl = wlo;
        bitpos = end_sc;
    }
    if (bitpos & 31) flush_ring(s.ring, obase, abase, lo, hi, wfl, wfl + 1);
    if (tid == 0) {
        const u64 bits = bitpos - (pos & 3ull) * 8ull;
        if (((bits + 7) >> 3) != (u64)cr.nbytes) atomicExch(err, 0xBAD00001u);
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------
// decoupled look-back: exclusive prefix of frame sizes in global frame order
// ----------------------------------------------------------------------------
constexpr u64 ST_AGG = 1ull << 62, ST_PRE = 2ull << 62, ST_MASK = (1ull << 62) - 1;

__device__ __forceinline__ u64 ld_status(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(u64 *p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// called by warp 0; returns the exclusive prefix in every lane
__device__ u64 lookback_exclusive(u64 *status, u32 g, u64 mine) {
    const int lane = threadIdx.x & 31;
    if (g == 0) {
        if (lane == 0) st_status(status, ST_PRE | mine);
        return 0;
    }
    if (lane == 0) st_status(status + g, ST_AGG | mine);
    u64 excl = 0;
    i64 idx = (i64)g - 1;
    for (;;) {
        const i64 j = idx - lane;
        u64 v = ST_PRE;                       // virtual predecessor before frame 0: prefix 0
        if (j >= 0) {
            do { v = ld_status(status + j); } while ((v >> 62) == 0);
        }
        const u32 pm = __ballot_sync(0xffffffffu, (v >> 62) == 2);
        u64 val = v & ST_MASK;
        if (pm) {
            const int first = __ffs(pm) - 1;  // nearest predecessor holding an inclusive prefix
            if (lane > first) val = 0;
            val = warp_sum64(val);
            excl += __shfl_sync(0xffffffffu, val, 0);
            break;
        }
        val = warp_sum64(val);
        excl += __shfl_sync(0xffffffffu, val, 0);
        idx -= 32;
    }
    if (lane == 0) st_status(status + g, ST_PRE | (excl + mine));
    return excl;
}

// ----------------------------------------------------------------------------
// the frame-encode kernel
// ----------------------------------------------------------------------------
// channel header bytes inside an ALPC frame, writer.rs:272-299 (none in a Raw-typed frame, :267-270)
__device__ __forceinline__ u32 chan_hdr_bytes(bool all_raw, const ChanResult &r) {
    if (all_raw) return 0;
    return 1u + 4u * (r.kind == 2 ? (u32)r.order : 0u) + 1u + 1u + (r.kind != 0 ? 1u : 0u);
}
__device__ __forceinline__ void put_u32le(uint8_t *p, u32 v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}

template <typename T>
__device__ void ingest_frame(Smem &s, const T *in, u32 len, u32 C, int16_t *planes, u32 stride, bool &loud_out,
                             i64 (&var)[3]) {
    const int tid = threadIdx.x;
    bool loud = false;
    i64 vl = 0, vr = 0, vs = 0;
    if (C == 2) {
        const u32 nf = len >> 1;
        for (u32 i = tid; i < nf; i += NT) {
            const float a = sample_f32<T>(in, 2 * (size_t)i), b = sample_f32<T>(in, 2 * (size_t)i + 1);
            loud |= is_loud(a) | is_loud(b);
            const i32 l = f32_to_i32(a), r = f32_to_i32(b);
            planes[i] = (int16_t)l;
            planes[stride + i] = (int16_t)r;
            const i32 sd = l - r;
            vl += (i64)l * l; vr += (i64)r * r; vs += (i64)sd * sd;     // encoder.rs:136-149
        }
        if ((len & 1) && tid == 0) {           // ragged tail: channel 0 gets one more sample (encoder.rs:84-90)
            const float a = sample_f32<T>(in, (size_t)len - 1);
            loud |= is_loud(a);
            planes[nf] = (int16_t)f32_to_i32(a);
        }
    } else if (C == 1) {
        for (u32 i = tid; i < len; i += NT) {
            const float a = sample_f32<T>(in, i);
            loud |= is_loud(a);
            planes[i] = (int16_t)f32_to_i32(a);
        }
    } else {
        for (u32 e = tid; e < len; e += NT) {
            const float a = sample_f32<T>(in, e);
            loud |= is_loud(a);
            const u32 c = e % C, i = e / C;
            planes[(size_t)c * stride + i] = (int16_t)f32_to_i32(a);
        }
    }
    // zero the padding behind each channel (chunk loads read up to the next multiple of 16)
    for (u32 c = 0; c < C; c++) {
        const u32 cl = len > c ? (len - c + C - 1) / C : 0;
        for (u32 i = cl + tid; i < stride; i += NT) planes[(size_t)c * stride + i] = 0;
    }
    loud_out = __syncthreads_or(loud) != 0;
    var[0] = vl; var[1] = vr; var[2] = vs;
    (void)s;
}

extern __shared__ __align__(16) unsigned char dyn_smem[];

__global__ void __launch_bounds__(NT, 1) k_encode_frames(const EncodeParams p) {
    Smem &s = *reinterpret_cast<Smem *>(dyn_smem);
    int16_t *smem_planes = reinterpret_cast<int16_t *>(dyn_smem + ((sizeof(Smem) + 15) & ~size_t(15)));
    const int tid = threadIdx.x;
    for (int i = tid; i < RING_WORDS; i += NT) s.ring[i] = 0;
    __syncthreads();
    ChanResult *cres = p.cres + (size_t)blockIdx.x * 256;
    const int level = p.level;
    const int P = order_of_level(level);
    const int fmax = P < 4 ? P : 4;
    const bool lpc_on = level >= 3 && P > 4;

    for (;;) {
        if (tid == 0) s.g = atomicAdd(p.ticket, 1u);
        __syncthreads();
        const u32 g = s.g;
        if (g >= p.n_frames) break;
        const uint2 fd = p.frames[g];
        const TrackDev tr = p.tracks[fd.x];
        const u32 C = tr.channels;
        const u64 spf_inter = (u64)tr.sample_rate * C;                     // encoder.rs:33, 53-58
        const u64 start = (u64)fd.y * spf_inter;
        const u64 end = min(start + spf_inter, tr.n_inter);
        const u32 len = (u32)(end - start);
        const u32 frame_samples = len / C;                                 // encoder.rs:67
        const u32 cl0 = (len + C - 1) / C;
        const u32 stride = (cl0 + 15u) & ~15u;
        int16_t *planes = ((u64)C * stride * 2 <= p.smem_plane_bytes)
                              ? smem_planes
                              : p.plane_scratch + (size_t)blockIdx.x * p.plane_scratch_elems;

        bool loud;
        i64 var[3];
        if (p.format == FLO_FMT_PCM16)
            ingest_frame<int16_t>(s, reinterpret_cast<const int16_t *>(tr.samples) + start, len, C, planes, stride, loud, var);
        else
            ingest_frame<float>(s, reinterpret_cast<const float *>(tr.samples) + start, len, C, planes, stride, loud, var);

        const u64 data_base = tr.static_off + FILE_HDR + 4ull + 20ull * tr.n_frames;   // writer.rs:51, 89-95

        if (!loud) {
            // Frame::silence, encoder.rs:70-76 / types.rs:221-229: type 0, C empty channels
            const u32 fsize = 6 + 4 * C;
            if (tid < 32) {
                u64 ex = lookback_exclusive(p.status, g, fsize);
                if (tid == 0) { s.frame_excl = ex; p.frame_excl[g] = ex; p.frame_size[g] = fsize; }
            }
            __syncthreads();
            uint8_t *o = p.out + data_base + s.frame_excl;
            if (tid == 0) { o[0] = 0; put_u32le(o + 1, frame_samples); o[5] = 0; }
            for (u32 i = tid; i < 4 * C; i += NT) o[6 + i] = 0;
            if (p.report) {
                for (u32 i = tid; i < REPORT_CH * NCAND; i += NT) {
                    flo_cand_report *r = p.report + (size_t)g * REPORT_CH * NCAND + i;
                    r->k = 0; r->pad = 0; r->size = -1;
                }
            }
            __syncthreads();
            continue;
        }

        // mid/side decision, encoder.rs:94-100, 131-153
        if (C == 2) {
            u64 v3[3] = {(u64)var[0], (u64)var[1], (u64)var[2]};
            block_sum64<3>(s, v3);
            if (tid == 0) {
                const i64 vl = (i64)s.tot64[0], vr = (i64)s.tot64[1], vs = (i64)s.tot64[2];
                s.ms = vs < (vl + vr) / 2 ? 1 : 0;
            }
            __syncthreads();
        } else if (tid == 0) {
            s.ms = 0;
        }
        __syncthreads();
        const int ms = s.ms;
        if (ms && (len & 1)) {                 // the unpaired tail sample of L is dropped by the zip (encoder.rs:160)
            if (tid == 0) planes[len >> 1] = 0;
            __syncthreads();
        }

        // per-channel predictor search
        bool all_raw = true;
        u32 fsize = 6;
        const u32 frame_type_alpc = (P >= 1 && P <= 12) ? (u32)P : 8u;     // FrameType::from_order, types.rs:69-85
        for (u32 c = 0; c < C; c++) {
            ChanView cv;
            u32 cl = len > c ? (len - c + C - 1) / C : 0;
            if (ms) cl = len >> 1;                                          // zip in to_mid_side truncates, encoder.rs:160-167
            cv.n = (int)cl;
            cv.msmode = ms ? (c == 0 ? 1 : 2) : 0;
            cv.pa = ms ? planes : planes + (size_t)c * stride;
            cv.pb = planes + stride;
            flo_cand_report *rep = (p.report && c < REPORT_CH) ? p.report + ((size_t)g * REPORT_CH + c) * NCAND : nullptr;
            if (cl == 0) {
                if (tid == 0) { cres[c].kind = 3; cres[c].order = 0; cres[c].k = 0; cres[c].nbytes = 0; }
                __syncthreads();
                continue;
            }
            const int Pa = (lpc_on && (int)cl > 5) ? P : 0;
            switch (Pa) {
                case 6: analyse_channel<6>(s, cv, level, fmax, rep); break;
                case 8: analyse_channel<8>(s, cv, level, fmax, rep); break;
                case 10: analyse_channel<10>(s, cv, level, fmax, rep); break;
                case 12: analyse_channel<12>(s, cv, level, fmax, rep); break;
                default: analyse_channel<0>(s, cv, level, fmax, rep); break;
            }
            __syncthreads();
            if (tid == 0) {
                // encode_channel_int, encoder.rs:184-216: strictly smaller wins, candidates in order
                i64 best = s.cand_size[0];
                int bj = 0;
                for (int j = 1; j < NCAND; j++) {
                    const i64 sz = s.cand_size[j];
                    if (sz >= 0 && sz < best) { best = sz; bj = j; }
                }
                ChanResult r;
                r.kind = bj == 0 ? 0 : (bj <= 5 ? 1 : 2);
                r.order = bj == 0 ? 0 : bj - 1;
                r.k = s.cand_k[bj];
                r.nbytes = (u32)best;
                for (int j = 0; j < MAXORD; j++) r.coef[j] = (r.kind == 2 && j < r.order) ? s.qc[r.order][j] : 0;
                r.shift = r.kind == 2 ? s.lpc_shift[r.order] : 0;
                cres[c] = r;
            }
            __syncthreads();
        }
        __threadfence_block();
        __syncthreads();

        // frame typing and size, encoder.rs:102-127, types.rs:242-267
        for (u32 c = 0; c < C; c++) {
            const ChanResult &r = cres[c];
            if (r.order > 0) all_raw = false;
        }
        for (u32 c = 0; c < C; c++) {
            const ChanResult &r = cres[c];
            fsize += 4 + chan_hdr_bytes(all_raw, r) + r.nbytes;
        }
        if (tid < 32) {
            u64 ex = lookback_exclusive(p.status, g, fsize);
            if (tid == 0) { s.frame_excl = ex; p.frame_excl[g] = ex; p.frame_size[g] = fsize; }
        }
        __syncthreads();

        // write the frame, writer.rs:236-301
        const u64 fpos = data_base + s.frame_excl;
        uint8_t *o = p.out;
        if (tid == 0) {
            o[fpos] = (uint8_t)(all_raw ? 254u : frame_type_alpc);
            put_u32le(o + fpos + 1, frame_samples);
            o[fpos + 5] = (uint8_t)(ms ? 1 : 0);
        }
        u64 pos = fpos + 6;
        for (u32 c = 0; c < C; c++) {
            const ChanResult r = cres[c];
            const int shift = r.shift;
            const u32 hdr = chan_hdr_bytes(all_raw, r);
            __syncthreads();
            if (tid < MAXORD) s.wcoef[tid] = r.coef[tid];
            __syncthreads();
            if (tid == 0) {
                put_u32le(o + pos, hdr + r.nbytes);
                if (!all_raw) {
                    uint8_t *h = o + pos + 4;
                    const u32 nco = r.kind == 2 ? (u32)r.order : 0u;
                    *h++ = (uint8_t)nco;
                    for (u32 j = 0; j < nco; j++) { put_u32le(h, (u32)r.coef[j]); h += 4; }
                    *h++ = (uint8_t)(r.kind == 2 ? shift : (r.kind == 1 ? 128 + r.order : 0));   // encoder.rs:243, 279
                    *h++ = (uint8_t)(r.kind == 0 ? 2 : 0);                                       // ResidualEncoding
                    if (r.kind != 0) *h++ = (uint8_t)r.k;
                }
            }
            if (r.kind != 3 && r.nbytes > 0) {
                ChanView cv;
                u32 cl = len > c ? (len - c + C - 1) / C : 0;
                if (ms) cl = len >> 1;
                cv.n = (int)cl;
                cv.msmode = ms ? (c == 0 ? 1 : 2) : 0;
                cv.pa = ms ? planes : planes + (size_t)c * stride;
                cv.pb = planes + stride;
                pack_channel(s, cv, r, shift, o, pos + 4 + hdr, p.err);
            }
            pos += 4 + hdr + r.nbytes;
        }
        if (tid == 0 && pos - fpos != fsize) atomicExch(p.err, 0xBAD00002u);
        __syncthreads();
    }
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

// CRC-32 of one 128-byte aligned-by-construction block given as bytes [a, a+128) of buf
__device__ __forceinline__ u32 crc_update_word(u32 c, u32 w, const u32 (*T)[256]) {
    c ^= w;
    return T[3][c & 0xff] ^ T[2][(c >> 8) & 0xff] ^ T[1][(c >> 16) & 0xff] ^ T[0][c >> 24];
}

// One warp per 64 KB segment of a track's DATA chunk (crc32.rs:23-30 restated as
// per-block CRCs combined with x^(8 len) shifts; result identical to the serial loop).
__global__ void __launch_bounds__(CRC_NT) k_crc_segments(const FinalParams p) {
    __shared__ u32 T[4][256];
    for (int i = threadIdx.x; i < 1024; i += CRC_NT) (&T[0][0])[i] = (&c_crc_slice[0][0])[i];
    __syncthreads();
    const u32 lane = threadIdx.x & 31;
    const u32 seg = (blockIdx.x * CRC_NT + threadIdx.x) >> 5;
    if (seg >= p.n_segs) return;
    u32 lo = 0, hi = p.n_tracks;
    while (hi - lo > 1) {
        const u32 mid = (lo + hi) >> 1;
        if (p.tracks[mid].first_seg <= seg) lo = mid; else hi = mid;
    }
    const TrackDev tr = p.tracks[lo];
    const u64 e0 = excl_at(p, tr.first_frame), e1 = excl_at(p, tr.first_frame + tr.n_frames);
    const u64 dsize = e1 - e0;
    const u64 soff = (u64)(seg - tr.first_seg) * CRC_SEG;
    if (soff >= dsize) return;
    const u64 d0 = tr.static_off + e0 + FILE_HDR + 4 + 20ull * tr.n_frames;
    const uint8_t *base = p.out + d0 + soff;
    const u32 slen = (u32)min((u64)CRC_SEG, dsize - soff);
    const u32 nblk = slen >> 7;                      // full 128-byte blocks
    const u32 mis = (u32)((uintptr_t)base & 3u);
    const u32 *wbase = reinterpret_cast<const u32 *>(base - mis);
    u32 crc = 0;                                     // crc("") = 0 is the identity of combine
    for (u32 b0 = 0; b0 < nblk; b0 += 32) {
        const u32 m = min(32u, nblk - b0);
        u32 c = 0;
        if (lane < m) {
            const u32 *w = wbase + (size_t)(b0 + lane) * 32;
            u32 r = 0xFFFFFFFFu;
            if (mis == 0) {
#pragma unroll 8
                for (int i = 0; i < 32; i++) r = crc_update_word(r, w[i], T);
            } else {
                u32 cur = w[0];
#pragma unroll 8
                for (int i = 0; i < 32; i++) {
                    const u32 nxt = w[i + 1];
                    r = crc_update_word(r, __funnelshift_r(cur, nxt, 8 * mis), T);
                    cur = nxt;
                }
            }
            c = ~r;
            c = multmodp(c_pow128[m - 1 - lane], c);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, o);
        crc = multmodp(c_pow128[m], crc) ^ c;
    }
    if (lane == 0) {
        u32 r = ~crc;
        for (u32 i = nblk << 7; i < slen; i++) r = (r >> 8) ^ T[0][(r ^ base[i]) & 0xff];
        p.seg_crc[seg] = ~r;
    }
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
    if (threadIdx.x == 0) {
        u32 crc = 0;
        const u64 nseg = (dsize + CRC_SEG - 1) / CRC_SEG;
        for (u64 j = 0; j < nseg; j++) {
            const u64 len = min((u64)CRC_SEG, dsize - j * CRC_SEG);
            const u32 mul = len == CRC_SEG ? c_powseg : x2nmodp(len, 3);
            crc = multmodp(mul, crc) ^ p.seg_crc[tr.first_seg + j];
        }
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
cudaError_t configure_encode_kernel(size_t dyn_smem) {
    return cudaFuncSetAttribute(k_encode_frames, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
}
cudaError_t launch_setup(const TrackDev *tracks, uint32_t n_tracks, uint2 *frames, uint32_t n_frames, cudaStream_t st) {
    if (n_frames == 0) return cudaSuccess;
    k_setup_frames<<<(n_frames + 255) / 256, 256, 0, st>>>(tracks, n_tracks, frames, n_frames);
    return cudaGetLastError();
}
cudaError_t launch_encode(const EncodeParams &p, int grid, size_t dyn_smem, cudaStream_t st) {
    if (p.n_frames == 0) return cudaSuccess;
    k_encode_frames<<<grid, NT, dyn_smem, st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_toc(const FinalParams &p, cudaStream_t st) {
    if (p.n_frames == 0) return cudaSuccess;
    k_write_toc<<<(p.n_frames + 255) / 256, 256, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_crc_segments(const FinalParams &p, cudaStream_t st) {
    if (p.n_segs == 0) return cudaSuccess;
    const uint32_t warps_per_block = CRC_NT / 32;
    k_crc_segments<<<(p.n_segs + warps_per_block - 1) / warps_per_block, CRC_NT, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_headers(const FinalParams &p, cudaStream_t st) {
    if (p.n_tracks == 0) return cudaSuccess;
    k_write_headers<<<p.n_tracks, 128, 0, st>>>(p);
    return cudaGetLastError();
}

}  // namespace flo
