// This file is the body of one build variant of the frame-encode kernel: it is included by
// flo_encode_nt{512,256,128}.cu inside namespace flo::FLO_VARIANT_NS with FLO_VARIANT_NT threads per CTA.
constexpr int NT = FLO_VARIANT_NT;
constexpr int NWARP = NT / 32;

// ----------------------------------------------------------------------------
// shared state of the frame-encode CTA
// ----------------------------------------------------------------------------
constexpr int GROUP = 2;              // channels analysed jointly (stereo = one group)
constexpr int NLPC = MAXORD - 4;      // LPC orders 5..12

// candidate states
constexpr int CS_ABSENT = 0, CS_EXACT = 1, CS_BOUNDED = 2;

struct ChanState {
    // layout of the coded channel (after the mid/side decision)
    const int16_t *pa, *pb;
    int msmode;                       // 0 plain, 1 mid = L + R, 2 side = L - R (encoder.rs:156-170)
    int n;                            // samples in this channel
    int nfull;                        // full chunks: n / CH
    int tail;                         // n % CH samples in the partial tail chunk
    // pass 1: fixed-predictor statistics + autocorrelation
    u64 fix_sum[5];
    u32 fix_or[5];
    i64 ac[MAXORD + 1];
    // Levinson-Durbin results, by order - 5
    double qd[NLPC][MAXORD];          // q / 2^shift as f64 (exact), for the FP64-pipe FIR
    i32 qc[NLPC][MAXORD];             // quantised coefficients (lpc.rs:263-273)
    i32 lpc_ok[NLPC];
    i32 lpc_shift[NLPC];
    i32 lpc_j0[NLPC];                 // guessed shift window {j0, j0 + 1} for sum(w >> j)
    double lpc_err[NLPC];             // prediction error after each order (window guess only)
    // pass 2: LPC statistics
    u64 l_sum[NLPC];
    u32 l_or[NLPC];
    u64 l_t0[NLPC], l_t1[NLPC];
    // candidates: 0 raw, 1..5 fixed 0..4, 6..13 lpc 5..12
    i32 cand_state[NCAND];
    i32 cand_k[NCAND];
    i64 cand_size[NCAND];             // exact bytes when CS_EXACT
    u64 cand_sumabs[NCAND];
    // exact pass (pass 3)
    i32 ex_cand;                      // candidate being evaluated this round (-1 none)
    u64 ex_s;
    u32 ex_max;
    u32 ex_fixed;                     // bit o set: fixed order o is evaluated this round (all open fixed ones at once)
    u64 ex_s5[5];
};

struct Smem {
    ChanState cs[GROUP];
    i32 wcoef[MAXORD];                // winner's coefficients while packing
    double wqd[MAXORD];
    u32 scan_warp[2][NWARP];
    u32 cnt[8];                       // analysis counters of this CTA (flushed at kernel end)
    u32 g;                            // current global frame
    i32 ms;                           // mid/side chosen (encoder.rs:94-100)
    i32 loud;
    u64 ms_var[3];
    u64 frame_excl;                   // exclusive prefix of frame sizes
    u32 ring[RING_WORDS];             // bit-packer staging ring (big-endian bit order words)
};

static size_t encode_static_smem() { return sizeof(Smem); }

__device__ __forceinline__ u64 warp_sum64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void atomic_add64(u64 *p, u64 v) { atomicAdd(reinterpret_cast<unsigned long long *>(p), v); }

// ----------------------------------------------------------------------------
// scalar pieces of the reference
// ----------------------------------------------------------------------------
// f32_to_i32, core/audio_constants.rs:18-20: (x * 32767.0).clamp(-32768, 32767) as i32.
// cvt.rzi.sat saturates and maps NaN to 0 exactly like Rust's `as i32`; saturating the
// truncated value equals truncating the clamped float because both bounds are integers.
__device__ __forceinline__ i32 f32_to_i32(float x) {
    // one conversion instruction: round toward zero, saturate to [-32768, 32767], NaN -> 0
    const float y = __fmul_rn(x, 32767.0f);
    short v;
    asm("cvt.rzi.sat.s16.f32 %0, %1;" : "=h"(v) : "f"(y));
    return (i32)v;
}
// silence test of encoder.rs:70: |x| < 1e-7 (NaN is not silent)
__device__ __forceinline__ bool is_loud(float x) { return !(fabsf(x) < 1e-7f); }
// reflo/src/audio.rs:247-254: s as f32 * (1.0 / 32768.0)
__device__ __forceinline__ float pcm_to_f32(int s) { return __fmul_rn((float)s, 1.0f / 32768.0f); }

template <typename T> __device__ __forceinline__ float sample_f32(const T *p, size_t i);
template <> __device__ __forceinline__ float sample_f32<float>(const float *p, size_t i) { return __ldg(p + i); }
template <> __device__ __forceinline__ float sample_f32<int16_t>(const int16_t *p, size_t i) { return pcm_to_f32(__ldg(p + i)); }

__device__ __forceinline__ int bitlen32(u32 v) { return 32 - __clz((int)v); }

// estimate_rice_parameter_i32, core/rice.rs:29-69, from OR(|r|) and sum(|r|):
// only the bit length of max|r| enters the rule, and OR has the same bit length as the maximum.
__device__ __forceinline__ int rice_k_or(u32 or_abs, u64 sum_abs, u32 n) {
    if (n == 0) return 4;
    if (or_abs == 0) return 0;
    const int bl = bitlen32(or_abs);              // bitlen(max_abs)
    const int min_k = bl >= 8 ? bl + 1 - 8 : 0;   // 2 max > 255  <=>  max >= 128; bits_needed = bl + 1
    const u32 mean = (sum_abs >> 32) == 0 ? (u32)sum_abs / n : (u32)(sum_abs / (u64)n);
    const int mean_k = mean > 0 ? bitlen32(mean) : 0;
    const int k = max(min_k, mean_k);
    return min(k, 15);
}

// lpc_order_from_level, encoder.rs:289-302
__device__ __forceinline__ int order_of_level(int level) {
    const int t[10] = {0, 2, 4, 4, 6, 8, 8, 10, 12, 12};
    return t[level < 0 ? 0 : (level > 9 ? 9 : level)];
}

// Size algebra (rice.rs:97-113): with u = zigzag(r), w = r ^ (r >> 31) = |r| - [r < 0]:
//   u >> k == w >> (k - 1) for k >= 1, and sum(u) = sum(w) + sum(|r|).
// So for S = sum(w >> max(k - 1, 0)):  bits = S + n (1 + k)  (k >= 1),  bits = S + sum|r| + n  (k == 0).
// The 255 cap of rice.rs:103 never binds: k >= bitlen(2 max|r|) - 8 makes u >> k <= 255.
__device__ __forceinline__ i64 rice_bytes(u64 S, u64 sum_abs, u32 n, int k) {
    const u64 bits = k >= 1 ? S + (u64)n * (u64)(1 + k) : S + sum_abs + (u64)n;
    return (i64)((bits + 7) >> 3);
}
// bounds on the encoded size from sum|r| and k alone: sum|r| / 2^j - n <= S <= sum|r| / 2^j, j = max(k - 1, 0)
__device__ __forceinline__ void rice_bounds(u64 sum_abs, u32 n, int k, i64 &lb, i64 &ub) {
    const int j = k >= 1 ? k - 1 : 0;
    const u64 hi = sum_abs >> j;
    const u64 c = (sum_abs + ((1ull << j) - 1)) >> j;
    const u64 lo = c > n ? c - n : 0;
    lb = rice_bytes(lo, sum_abs, n, k);
    ub = rice_bytes(hi, sum_abs, n, k);
}

// levinson_durbin_int, lpc.rs:225-276 -- sequential f64, every product and sum rounded
// separately.  The recursion is prefix consistent (the order-m result is the state after
// iteration m-1), so one run to order P yields every order 5..P.  Also derives, per order,
// the guessed shift window for the single-pass size evaluation from the prediction error.
template <int P>
__device__ void levinson_all_orders(ChanState &cs) {      // called by a full warp
    const int lane = threadIdx.x & 31;
    if (lane < NLPC) { cs.lpc_ok[lane] = 0; cs.lpc_shift[lane] = 0; cs.lpc_j0[lane] = 0; }
    __syncwarp();
    // (1) the recursion itself is sequential: lane 0.  Unquantised coefficients of every order >= 5 are
    //     parked in cs.qd (overwritten by their quantised form below), the prediction error in cs.lpc_err.
    if (lane == 0 && cs.ac[0] != 0) {
        double a[P > 0 ? P : 1], nc[P > 0 ? P : 1], acd[P + 1];
#pragma unroll
        for (int i = 0; i < P; i++) a[i] = 0.0;
#pragma unroll
        for (int i = 0; i <= P; i++) acd[i] = (double)cs.ac[i];
        double err = acd[0];
        bool alive = true;
#pragma unroll
        for (int i = 0; i < P; i++) {
            if (!alive) break;
            double lambda = acd[i + 1];
#pragma unroll
            for (int j = 0; j < i; j++) lambda = __dsub_rn(lambda, __dmul_rn(a[j], acd[i - j]));
            if (fabs(err) < 1e-10) { alive = false; break; }
            const double gamma = __ddiv_rn(lambda, err);
            if (fabs(gamma) >= 1.0) { alive = false; break; }
            nc[i] = gamma;
#pragma unroll
            for (int j = 0; j < i; j++) nc[j] = __dsub_rn(a[j], __dmul_rn(gamma, a[i - 1 - j]));
#pragma unroll
            for (int j = 0; j <= i; j++) a[j] = nc[j];
            err = __dmul_rn(err, __dsub_rn(1.0, __dmul_rn(gamma, gamma)));
            if (i + 1 >= 5) {
#pragma unroll
                for (int j = 0; j <= i; j++) cs.qd[i + 1 - 5][j] = a[j];
                cs.lpc_err[i + 1 - 5] = err;
                cs.lpc_ok[i + 1 - 5] = 2;                  // reached; validated in (2)
            }
        }
    }
    __syncwarp();
    // (2) per order (lane t = order - 5): max |a|, shift, and the size-window guess
    if (lane < P - 4 && cs.lpc_ok[lane] == 2) {
        const int o = 5 + lane;
        double mx = 0.0;
        for (int j = 0; j < o; j++) { const double t = fabs(cs.qd[lane][j]); if (t == t && t > mx) mx = t; }
        int ok = 0;
        if (!(mx == 0.0 || isinf(mx))) {
            // shift = min(floor(log2(2^30 / max)) as u8, 15); floor(log2(v)) of a positive finite
            // double is its binary exponent (|a_j| <= C(12,6) = 924 makes this 15 in practice).
            const double v = __ddiv_rn(1073741824.0, mx);
            const int e = isinf(v) ? 255 : ilogb(v);
            cs.lpc_shift[lane] = e < 0 ? 0 : (e > 15 ? 15 : e);
            ok = 1;
            // Heuristic only (exactness never depends on it): mean|r| ~ 0.64 * rms(r), rms^2 ~ err / n.
            // The window {j0, j0+1} must contain max(k-1, 0); a miss is re-evaluated exactly in pass 3.
            // log2 via the exponent and a linear mantissa term is accurate to 0.09, ample here.
            const double err = cs.lpc_err[lane];
            const double rms2 = err > 0.0 ? err / (double)cs.n : 0.0;
            double lg = -10.0;
            if (rms2 > 1e-30) {
                int ex;
                const double m = frexp(rms2, &ex);        // rms2 = m 2^ex, m in [0.5, 1)
                lg = 0.5 * ((double)ex + 2.0 * m - 2.0) - 0.64;
            }
            const int j0 = (int)floor(lg - 0.5);
            cs.lpc_j0[lane] = j0 < 0 ? 0 : (j0 > 14 ? 14 : j0);
        }
        cs.lpc_ok[lane] = ok;
    }
    __syncwarp();
    // (3) quantise every (order, j) pair in parallel (lpc.rs:263-273)
    for (int item = lane; item < (P - 4) * P; item += 32) {
        const int t = item / (P > 0 ? P : 1), j = item % (P > 0 ? P : 1);
        if (j < 5 + t && cs.lpc_ok[t] == 1) {
            const int shift = cs.lpc_shift[t];
            const double scale = (double)(1 << shift), inv_scale = 1.0 / scale;      // powers of two: exact
            // f64::round (half away from zero): trunc, then one more if the (exact) remainder reaches 1/2
            const double y = __dmul_rn(cs.qd[t][j], scale);
            double q = trunc(y);
            if (fabs(y - q) >= 0.5) q += copysign(1.0, y);
            const i32 qi = q >= 2147483647.0 ? 2147483647 : (q <= -2147483648.0 ? (-2147483647 - 1) : (i32)q);
            cs.qc[t][j] = qi;
            cs.qd[t][j] = (double)qi * inv_scale;
        }
    }
    __syncwarp();
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
// 16-bit pair -> two i32 through the integer dot-product unit (off the ALU pipe): dp2a_lo(w, b, c) =
// c + w.lo * b.byte0 + w.hi * b.byte1, so b = 0x0001 picks the low half, 0x0100 the high half and
// 0x00ff / 0xff00 subtract them.  MODE 0: plain, 1: accumulate (+), 2: accumulate (-).
template <int MODE>
__device__ __forceinline__ void unpack8_dp(const int4 v, i32 *t) {
    const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (MODE == 0) { t[2 * i] = __dp2a_lo(w[i], 0x0001, 0); t[2 * i + 1] = __dp2a_lo(w[i], 0x0100, 0); }
        if (MODE == 1) { t[2 * i] = __dp2a_lo(w[i], 0x0001, t[2 * i]); t[2 * i + 1] = __dp2a_lo(w[i], 0x0100, t[2 * i + 1]); }
        if (MODE == 2) { t[2 * i] = __dp2a_lo(w[i], 0x00ff, t[2 * i]); t[2 * i + 1] = __dp2a_lo(w[i], 0xff00, t[2 * i + 1]); }
    }
}
template <int NH, int MODE>
__device__ __forceinline__ void load_plane(const int16_t *pl, int i0, i32 (&x)[NH + CH]) {
    static_assert(CH == 8 || CH == 16, "chunk of 8 or 16 samples");
    // x[NH - h .. NH): history, x[NH .. NH + CH): the chunk; 8-sample pieces that reach before the frame are zero
    const int4 *p = reinterpret_cast<const int4 *>(pl + i0);
    const int4 z = make_int4(0, 0, 0, 0);
    i32 t[16 + CH];
    if (MODE != 0) {
#pragma unroll
        for (int i = 0; i < 16 + CH; i++) t[i] = (i >= 16 - NH) ? x[i - (16 - NH)] : 0;
    }
    unpack8_dp<MODE>(p[0], t + 16);
    if (CH == 16) unpack8_dp<MODE>(p[1], t + 24);
    if (NH > 8) unpack8_dp<MODE>(i0 >= 16 ? p[-2] : z, t);
    if (NH > 0) unpack8_dp<MODE>(i0 >= 8 ? p[-1] : z, t + 8);
#pragma unroll
    for (int i = 0; i < NH + CH; i++) x[i] = t[16 - NH + i];
}
template <int NH>
__device__ __forceinline__ void load_x(const ChanState &cs, int i0, i32 (&x)[NH + CH]) {
    load_plane<NH, 0>(cs.pa, i0, x);
    const int msmode = cs.msmode;                      // mid = L + R, side = L - R (encoder.rs:156-170)
    if (msmode == 1) load_plane<NH, 1>(cs.pb, i0, x);
    else if (msmode == 2) load_plane<NH, 2>(cs.pb, i0, x);
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
__device__ __forceinline__ void fixed_chunk(const i32 *x /* 4 history + CH */, bool first, F &&fn) {
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

// int -> f64 on the conversion pipe.  `volatile` pins each conversion where it is written, so the
// compiler neither keeps a whole chunk of doubles alive (spills) nor re-converts a sample per use.
__device__ __forceinline__ double cvt_f64(i32 x) {
    double d;
    asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d) : "r"(x));
    return d;
}

// calc_residuals_int, lpc.rs:279-298, on the FP64 pipe.  With c[t] = q[t] / 2^shift (exact) the
// chain of fused multiply-adds holds sum(q[t] * s[i-1-t]) / 2^shift exactly (|sum q s| < 2^53),
// so floor() of it equals the reference's arithmetic `pred >> shift`; adding 1.5 * 2^52 with
// round-down leaves that floor, modulo 2^32, in the low word -- the `pred as i32` truncation.
// x: H history + CH samples.  Only a sliding window of O converted samples is alive at a time.
template <int O, int H, class F>
__device__ __forceinline__ void lpc_chunk(const i32 *x, bool first, const double *qd, F &&fn) {
    double c[O];
#pragma unroll
    for (int t = 0; t < O; t++) c[t] = qd[t];
    double w[O + CH];                                  // w[i] = f64(x[H - O + i])
#pragma unroll
    for (int t = 0; t < O; t++) w[t] = cvt_f64(x[H - O + t]);
#pragma unroll
    for (int j = 0; j < CH; j++) {
        double p = 0.0;
#pragma unroll
        for (int t = 0; t < O; t++) p = __fma_rn(c[t], w[O + j - 1 - t], p);
        const int pi = __double2loint(__dadd_rd(p, 6755399441055744.0));
        i32 r = (i32)((u32)x[H + j] - (u32)pi);
        if (j < O) r = first ? x[H + j] : r;           // warm-up, lpc.rs:283-285
        fn(j, r);
        if (j + 1 < CH) w[O + j] = cvt_f64(x[H + j]);
    }
}

// ----------------------------------------------------------------------------
// scalar residuals at one sample index (used for the < 16 samples behind the last full chunk)
// ----------------------------------------------------------------------------
__device__ __forceinline__ i32 sample_at(const ChanState &cs, int i) {
    if (i < 0 || i >= cs.n) return 0;
    const i32 a = cs.pa[i];
    if (cs.msmode == 0) return a;
    const i32 b = cs.pb[i];
    return cs.msmode == 1 ? a + b : a - b;
}
// fixed_predictor_residuals (lpc.rs:301-359) at index i: the min(o, i)-th finite difference
__device__ __noinline__ i32 fixed_residual_at(const ChanState &cs, int o, int i) {
    const int oo = o < i ? o : i;
    const i32 binom[5][5] = {{1, 0, 0, 0, 0}, {1, -1, 0, 0, 0}, {1, -2, 1, 0, 0}, {1, -3, 3, -1, 0}, {1, -4, 6, -4, 1}};
    u32 r = 0;
    for (int t = 0; t <= oo; t++) r += (u32)binom[oo][t] * (u32)sample_at(cs, i - t);
    return (i32)r;
}
// calc_residuals_int (lpc.rs:279-298) at index i, in the reference's own i64 arithmetic
__device__ __noinline__ i32 lpc_residual_at(const ChanState &cs, int o, int i) {
    const i32 x = sample_at(cs, i);
    if (i < o) return x;
    i64 pred = 0;
    for (int t = 0; t < o; t++) pred += (i64)cs.qc[o - 5][t] * (i64)sample_at(cs, i - 1 - t);
    pred >>= cs.lpc_shift[o - 5];
    return (i32)((u32)x - (u32)(i32)pred);
}

// ----------------------------------------------------------------------------
// Chunk loop of one analysis pass.  With two channels in the group the even warps take channel
// 0 and the odd warps channel 1, so a thread accumulates for one channel only and flushes once
// (per 64 rounds: the 32-bit partial sums hold at least 64 chunks).  The unrolled body sees full
// chunks only; the < 16 samples behind the last full chunk are added by the channel's last warp,
// one lane per sample, through the scalar functions above.
// ----------------------------------------------------------------------------
template <class Reset, class Body, class Tail, class Flush>
__device__ __forceinline__ void for_chunks(const Smem &s, int nch, Reset &&reset, Body &&body, Tail &&tailfn, Flush &&flush) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = nch == 2 ? (wid & 1) : 0;
    const int wi = nch == 2 ? (wid >> 1) : wid;
    const int nthr = nch == 2 ? NT / 2 : NT;
    const int ti = wi * 32 + lane;
    const bool tail_warp = wi == nthr / 32 - 1;
    const int nfull = s.cs[c].nfull;
    int base = 0;
    do {
        reset();
        const int end = min(nfull, base + nthr * 64);
        for (int chunk = base + ti; chunk < end; chunk += nthr) body(c, chunk);
        if (end == nfull && tail_warp && lane < s.cs[c].tail) tailfn(c, nfull * CH + lane);
        flush(c);
        base = end;
    } while (base < nfull);
}

// ---- pass 1: fixed-predictor statistics (sum|r|, OR|r|) + autocorrelation (lpc.rs:213-221) ----
template <int P>
__device__ void pass1(Smem &s, int nch) {
    constexpr int NH = P > 4 ? P : 4;
    u32 fsum[5], forr[5];               // |r| <= 2^20 for the fixed predictors: 2^24 per chunk
    double acc[P + 1];
    const int lane = threadIdx.x & 31;
    for_chunks(
        s, nch,
        [&]() {
#pragma unroll
            for (int o = 0; o < 5; o++) { fsum[o] = 0; forr[o] = 0; }
#pragma unroll
            for (int l = 0; l <= P; l++) acc[l] = 0.0;
        },
        [&](int c, int chunk) {
            const ChanState &cs = s.cs[c];
            const int i0 = chunk * CH;
            i32 x[NH + CH];
            load_x<NH>(cs, i0, x);
            if constexpr (P > 0) {
                // exact: |x| <= 2^16, so every partial sum is an integer far below 2^53
                double w[P + CH];
#pragma unroll
                for (int t = 0; t < P; t++) w[t] = cvt_f64(x[NH - P + t]);
#pragma unroll
                for (int j = 0; j < CH; j++) {
                    w[P + j] = cvt_f64(x[NH + j]);
#pragma unroll
                    for (int l = 0; l <= P; l++) acc[l] = __fma_rn(w[P + j], w[P + j - l], acc[l]);
                }
            }
            // samples are accumulated in pairs: one 3-input add (IADD3) and one 3-input OR (LOP3) per two samples
            u32 p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0;
            fixed_chunk(x + (NH - 4), i0 == 0, [&](int j, i32 r0, i32 r1, i32 r2, i32 r3, i32 r4) {
                const u32 a0 = (u32)abs(r0), a1 = (u32)abs(r1), a2 = (u32)abs(r2), a3 = (u32)abs(r3), a4 = (u32)abs(r4);
                if (j & 1) {
                    fsum[0] += p0 + a0; fsum[1] += p1 + a1; fsum[2] += p2 + a2; fsum[3] += p3 + a3; fsum[4] += p4 + a4;
                    forr[0] |= p0 | a0; forr[1] |= p1 | a1; forr[2] |= p2 | a2; forr[3] |= p3 | a3; forr[4] |= p4 | a4;
                } else {
                    p0 = a0; p1 = a1; p2 = a2; p3 = a3; p4 = a4;
                }
            });
        },
        [&](int c, int i) {
            const ChanState &cs = s.cs[c];
#pragma unroll
            for (int o = 0; o < 5; o++) {
                const u32 a = (u32)abs(fixed_residual_at(cs, o, i));
                fsum[o] += a; forr[o] |= a;
            }
            if constexpr (P > 0) {
                const double xi = (double)sample_at(cs, i);
#pragma unroll
                for (int l = 0; l <= P; l++) acc[l] = __fma_rn(xi, (double)sample_at(cs, i - l), acc[l]);
            }
        },
        [&](int c) {
            ChanState &cs = s.cs[c];
#pragma unroll
            for (int o = 0; o < 5; o++) {
                const u64 t = warp_sum64((u64)fsum[o]);
                const u32 r = __reduce_or_sync(0xffffffffu, forr[o]);
                if (lane == 0) { atomic_add64(&cs.fix_sum[o], t); atomicOr(&cs.fix_or[o], r); }
            }
            if constexpr (P > 0) {
#pragma unroll
                for (int l = 0; l <= P; l++) {
                    const u64 t = warp_sum64((u64)__double2ll_rn(acc[l]));
                    if (lane == 0) atomic_add64(reinterpret_cast<u64 *>(&cs.ac[l]), t);
                }
            }
        });
}

// ---- pass 2: LPC candidates 5..P: sum|r|, OR|r| and sum(w >> j) for the guessed window ----
template <int P>
__device__ void pass2(Smem &s, int nch) {
    constexpr int NO = P - 4;
    // 32-bit partial sums: only candidates with OR|r| < 2^21 are ever used (after_pass2), 2^25 per chunk
    u32 lsum[NO], lt0[NO], lt1[NO], lorr[NO];
    const int lane = threadIdx.x & 31;
    for_chunks(
        s, nch,
        [&]() {
#pragma unroll
            for (int i = 0; i < NO; i++) { lsum[i] = 0; lt0[i] = 0; lt1[i] = 0; lorr[i] = 0; }
        },
        [&](int c, int chunk) {
            const ChanState &cs = s.cs[c];
            const int i0 = chunk * CH;
            i32 x[P + CH];
            load_x<P>(cs, i0, x);
            ForOrders<5, P>::run([&](auto oc) {
                constexpr int O = decltype(oc)::value;
                if (cs.lpc_ok[O - 5]) {
                    const int j0 = cs.lpc_j0[O - 5];
                    u32 sa = lsum[O - 5], t0 = lt0[O - 5], t1 = lt1[O - 5], orr = lorr[O - 5];
                    u32 pa = 0, pw = 0, pv = 0;
                    lpc_chunk<O, P>(x, i0 == 0, cs.qd[O - 5], [&](int j, i32 r) {
                        const u32 a = (u32)abs(r);
                        const u32 ws = (a + (u32)(r >> 31)) >> j0;     // w = |r| - [r < 0]
                        const u32 wv = ws >> 1;
                        if (j & 1) {                               // pairs: 3-input add / OR per two samples
                            sa += pa + a; orr |= pa | a;
                            t0 += pw + ws; t1 += pv + wv;
                        } else {
                            pa = a; pw = ws; pv = wv;
                        }
                    });
                    lsum[O - 5] = sa; lt0[O - 5] = t0; lt1[O - 5] = t1; lorr[O - 5] = orr;
                }
            });
        },
        [&](int c, int i) {
            const ChanState &cs = s.cs[c];
            ForOrders<5, P>::run([&](auto oc) {
                constexpr int O = decltype(oc)::value;
                if (cs.lpc_ok[O - 5]) {
                    const i32 r = lpc_residual_at(cs, O, i);
                    const u32 a = (u32)abs(r);
                    const u32 ws = (a + (u32)(r >> 31)) >> cs.lpc_j0[O - 5];
                    lsum[O - 5] += a; lorr[O - 5] |= a; lt0[O - 5] += ws; lt1[O - 5] += ws >> 1;
                }
            });
        },
        [&](int c) {
            ChanState &cs = s.cs[c];
#pragma unroll
            for (int i = 0; i < NO; i++) {
                const u64 a = warp_sum64((u64)lsum[i]), b = warp_sum64((u64)lt0[i]), d = warp_sum64((u64)lt1[i]);
                const u32 r = __reduce_or_sync(0xffffffffu, lorr[i]);
                if (lane == 0) {
                    atomic_add64(&cs.l_sum[i], a); atomic_add64(&cs.l_t0[i], b); atomic_add64(&cs.l_t1[i], d);
                    atomicOr(&cs.l_or[i], r);
                }
            }
        });
}

// residuals of one chunk for candidate MODE (0..4 fixed, 5..12 LPC, 13 raw samples)
template <int MODE, class F>
__device__ __forceinline__ void cand_chunk(const ChanState &cs, int i0, const double *qd, F &&fn) {
    if constexpr (MODE == 13) {
        i32 x[CH];
        load_x<0>(cs, i0, x);
#pragma unroll
        for (int j = 0; j < CH; j++) fn(j, x[j]);
    } else if constexpr (MODE <= 4) {
        i32 xf[4 + CH];
        load_x<4>(cs, i0, xf);
        fixed_chunk(xf, i0 == 0, [&](int j, i32 r0, i32 r1, i32 r2, i32 r3, i32 r4) {
            fn(j, MODE == 0 ? r0 : MODE == 1 ? r1 : MODE == 2 ? r2 : MODE == 3 ? r3 : r4);
        });
    } else {
        constexpr int H = MODE <= 8 ? 8 : 12;
        i32 x[H + CH];
        load_x<H>(cs, i0, x);
        lpc_chunk<MODE, H>(x, i0 == 0, qd, fn);
    }
}

// ---- pass 3: exact max|r| and S = sum(w >> j) for one still-open candidate per channel ----
template <int P>
__device__ void pass3(Smem &s, int nch) {
    u64 S;
    u32 mx;
    u32 S5[5];                         // 32-bit partial sums: (w >> j) < 2^21 per sample, flushed every 64 chunks
    const int lane = threadIdx.x & 31;
    for_chunks(
        s, nch, [&]() { S = 0; mx = 0; S5[0] = S5[1] = S5[2] = S5[3] = S5[4] = 0; },
        [&](int c, int chunk) {
            const ChanState &cs = s.cs[c];
            const int cand = cs.ex_cand;
            if (cand < 0) return;
            const int i0 = chunk * CH;
            if (cs.ex_fixed) {
                // all open fixed candidates share one difference chain (lpc.rs:301-359); S_o for unevaluated
                // orders is computed too and simply not used
                const int j0 = max(cs.cand_k[1] - 1, 0), j1 = max(cs.cand_k[2] - 1, 0), j2 = max(cs.cand_k[3] - 1, 0),
                          j3 = max(cs.cand_k[4] - 1, 0), j4 = max(cs.cand_k[5] - 1, 0);
                i32 xf[4 + CH];
                load_x<4>(cs, i0, xf);
                fixed_chunk(xf, i0 == 0, [&](int j, i32 r0, i32 r1, i32 r2, i32 r3, i32 r4) {
                    S5[0] += ((u32)abs(r0) + (u32)(r0 >> 31)) >> j0; S5[1] += ((u32)abs(r1) + (u32)(r1 >> 31)) >> j1;
                    S5[2] += ((u32)abs(r2) + (u32)(r2 >> 31)) >> j2; S5[3] += ((u32)abs(r3) + (u32)(r3 >> 31)) >> j3;
                    S5[4] += ((u32)abs(r4) + (u32)(r4 >> 31)) >> j4;
                });
                return;
            }
            const int k = cs.cand_k[cand];
            const int jj = k >= 1 ? k - 1 : 0;
            u64 acc = 0;
            auto fn = [&](int j, i32 r) {
                const u32 a = (u32)abs(r);
                mx = max(mx, a);
                acc += (a + (u32)(r >> 31)) >> jj;
            };
            const int mode = cand - 1;                 // fixed 0..4 -> 0..4, lpc 5..12 -> 5..12
            const double *qd = mode >= 5 ? cs.qd[mode - 5] : nullptr;
            switch (mode) {
                case 0: cand_chunk<0>(cs, i0, qd, fn); break;
                case 1: cand_chunk<1>(cs, i0, qd, fn); break;
                case 2: cand_chunk<2>(cs, i0, qd, fn); break;
                case 3: cand_chunk<3>(cs, i0, qd, fn); break;
                case 4: cand_chunk<4>(cs, i0, qd, fn); break;
                default:
                    if constexpr (P > 0) {
                        ForOrders<5, P>::run([&](auto oc) {
                            constexpr int O = decltype(oc)::value;
                            if (mode == O) cand_chunk<O>(cs, i0, qd, fn);
                        });
                    }
                    break;
            }
            S += acc;
        },
        [&](int c, int i) {
            const ChanState &cs = s.cs[c];
            const int cand = cs.ex_cand;
            if (cand < 0) return;
            if (cs.ex_fixed) {
#pragma unroll
                for (int o = 0; o < 5; o++) {
                    const i32 r = fixed_residual_at(cs, o, i);
                    S5[o] += ((u32)abs(r) + (u32)(r >> 31)) >> max(cs.cand_k[1 + o] - 1, 0);
                }
                return;
            }
            const int k = cs.cand_k[cand];
            const int jj = k >= 1 ? k - 1 : 0;
            const int mode = cand - 1;
            const i32 r = mode <= 4 ? fixed_residual_at(cs, mode, i) : lpc_residual_at(cs, mode, i);
            const u32 a = (u32)abs(r);
            mx = max(mx, a);
            S += (a + (u32)(r >> 31)) >> jj;
        },
        [&](int c) {
            ChanState &cs = s.cs[c];
            const u64 t = warp_sum64(S);
            const u32 m = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0 && cs.ex_cand >= 0) { atomic_add64(&cs.ex_s, t); atomicMax(&cs.ex_max, m); }
            if (cs.ex_fixed) {
#pragma unroll
                for (int o = 0; o < 5; o++) {
                    const u64 t5 = warp_sum64((u64)S5[o]);
                    if (lane == 0) atomic_add64(&cs.ex_s5[o], t5);
                }
            }
        });
}

// ----------------------------------------------------------------------------
// candidate bookkeeping (one thread per channel)
// ----------------------------------------------------------------------------
// The candidate bookkeeping below runs on one full warp per channel: lane j owns candidate j
// (0 raw, 1..5 fixed 0..4, 6..13 LPC 5..12); lane 0 additionally runs the Levinson recursion.

// after pass 1: k of every fixed candidate, raw size; then Levinson
template <int P>
__device__ void after_pass1_warp(ChanState &cs, int fmax, bool lpc_on) {
    const int lane = threadIdx.x & 31;
    const u32 n = (u32)cs.n;
    if (lane < NCAND) {
        int state = CS_ABSENT, k = 0;
        i64 size = -1;
        u64 sumabs = 0;
        if (lane == 0) { state = CS_EXACT; size = 2ll * n; }                       // encode_raw, encoder.rs:220-226
        if (lane >= 1 && lane <= 1 + fmax) {
            const int o = lane - 1;
            state = CS_BOUNDED;
            k = rice_k_or(cs.fix_or[o], cs.fix_sum[o], n);
            sumabs = cs.fix_sum[o];
        }
        cs.cand_state[lane] = state; cs.cand_k[lane] = k; cs.cand_size[lane] = size; cs.cand_sumabs[lane] = sumabs;
    }
    if (lane < NLPC) cs.lpc_ok[lane] = 0;
    __syncwarp();
    if (lpc_on && cs.n > 5) {
        if constexpr (P > 0) levinson_all_orders<P>(cs);
        if (lane < NLPC && cs.n <= 5 + lane) cs.lpc_ok[lane] = 0;                 // encoder.rs:255-257
    }
    __syncwarp();
}

// after pass 2 (lanes 6..13): resolve the LPC candidates (encoder.rs:262-286)
__device__ void after_pass2_warp(ChanState &cs, int P, u32 *counters) {
    const int lane = threadIdx.x & 31;
    const u32 n = (u32)cs.n;
    bool hit = false, miss = false;
    const int o = lane - 1;
    if (lane >= 6 && o <= P && cs.lpc_ok[o - 5]) {
        const int i = o - 5;
        const u32 orr = cs.l_or[i];
        const int bl = bitlen32(orr);
        if (bl < 21) {                                                            // else max|r| >= 2^20 > 1_000_000: rejected
            const int k = rice_k_or(orr, cs.l_sum[i], n);
            cs.cand_k[lane] = k;
            cs.cand_sumabs[lane] = cs.l_sum[i];
            const int jj = k >= 1 ? k - 1 : 0;
            if (bl <= 19 && (jj == cs.lpc_j0[i] || jj == cs.lpc_j0[i] + 1)) {
                const u64 S = jj == cs.lpc_j0[i] ? cs.l_t0[i] : cs.l_t1[i];
                cs.cand_state[lane] = CS_EXACT;
                cs.cand_size[lane] = rice_bytes(S, cs.l_sum[i], n, k);
                hit = true;
            } else {
                cs.cand_state[lane] = CS_BOUNDED;                                 // window miss or 2^19 <= max|r| < 2^20
                miss = true;
            }
        }
    }
    const u32 hm = __ballot_sync(0xffffffffu, hit), mm = __ballot_sync(0xffffffffu, miss);
    if (lane == 0) {
        if (hm) atomicAdd(counters + 2, (u32)__popc(hm));
        if (mm) atomicAdd(counters + 3, (u32)__popc(mm));
    }
    __syncwarp();
}

__device__ __forceinline__ u64 warp_min64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const u64 t = __shfl_xor_sync(0xffffffffu, v, o); v = t < v ? t : v; }
    return v;
}

// Pick the next candidate that still needs an exact evaluation: bounded, and its lower bound
// does not exceed the best upper bound (otherwise it can never be the strictly-smallest one).
// Among those the one with the smallest lower bound goes first (it tightens the bound most).
__device__ int next_open_candidate_warp(ChanState &cs, bool prune, u32 *counters) {
    const int lane = threadIdx.x & 31;
    const u32 n = (u32)cs.n;
    int state = CS_ABSENT;
    i64 lb = 0, ub = 0;
    u64 ub_eff = ~0ull;
    if (lane < NCAND) {
        state = cs.cand_state[lane];
        if (state == CS_EXACT) ub_eff = (u64)cs.cand_size[lane];
        else if (state == CS_BOUNDED) {
            rice_bounds(cs.cand_sumabs[lane], n, cs.cand_k[lane], lb, ub);
            // an LPC candidate with 2^19 <= max|r| < 2^20 may still be rejected: its upper bound does not count
            const bool maybe_rejected = lane >= 6 && bitlen32(cs.l_or[lane - 6]) == 20;
            if (!maybe_rejected) ub_eff = (u64)ub;
        }
    }
    const u64 best_ub = warp_min64(ub_eff);
    const bool dead = state == CS_BOUNDED && prune && (u64)lb > best_ub;          // provably not the winner
    if (dead) { cs.cand_state[lane] = CS_ABSENT; state = CS_ABSENT; }
    const u64 key = state == CS_BOUNDED ? (((u64)lb << 8) | (u64)lane) : ~0ull;   // smallest lb, then lowest index
    const u64 best = warp_min64(key);
    const int pick = best == ~0ull ? -1 : (int)(best & 0xff);
    const u32 dm = __ballot_sync(0xffffffffu, dead);
    // when a fixed predictor is due, every still-open fixed predictor of the channel is evaluated in the same round
    u32 fm = (pick >= 1 && pick <= 5) ? ((__ballot_sync(0xffffffffu, state == CS_BOUNDED) >> 1) & 0x1fu) : 0u;
    const u32 nfm = (u32)__popc(fm);
    if (nfm == 1) fm = 0;                       // a single one: the one-candidate path is cheaper
    if (lane == 0) {
        if (dm) atomicAdd(counters + 5, (u32)__popc(dm));
        if (nfm) atomicAdd(counters + 4, nfm);
        cs.ex_cand = pick; cs.ex_s = 0; cs.ex_max = 0;
        cs.ex_fixed = fm;
        for (int o = 0; o < 5; o++) cs.ex_s5[o] = 0;
    }
    __syncwarp();
    return pick;
}

__device__ void after_pass3(ChanState &cs) {
    const int c = cs.ex_cand;
    if (c < 0) return;
    const u32 n = (u32)cs.n;
    if (cs.ex_fixed) {
        for (int o = 0; o < 5; o++)
            if (cs.ex_fixed >> o & 1) {
                cs.cand_state[1 + o] = CS_EXACT;
                cs.cand_size[1 + o] = rice_bytes(cs.ex_s5[o], cs.cand_sumabs[1 + o], n, cs.cand_k[1 + o]);
            }
        return;
    }
    if (c >= 6 && cs.ex_max > 1000000u) { cs.cand_state[c] = CS_ABSENT; return; }   // encoder.rs:269-272
    cs.cand_state[c] = CS_EXACT;
    cs.cand_size[c] = rice_bytes(cs.ex_s, cs.cand_sumabs[c], n, cs.cand_k[c]);
}

// ----------------------------------------------------------------------------
// bit packer
// ----------------------------------------------------------------------------
// Codes of one chunk for the winner: zigzag(r) (rice.rs:96), or the two little-endian bytes of
// the sample as one 16-bit MSB-first code for a raw channel ((s as i16).to_le_bytes(), encoder.rs:222-224).
template <int MODE>
__device__ __forceinline__ void chunk_codes(const ChanState &cs, int i0, const double *qd, u32 (&u)[CH]) {
    cand_chunk<MODE>(cs, i0, qd, [&](int j, i32 r) {
        if constexpr (MODE == 13) {
            const u32 v = (u32)r & 0xffffu;
            u[j] = ((v & 0xff) << 8) | (v >> 8);
        } else {
            u[j] = ((u32)r << 1) ^ (u32)(r >> 31);
        }
    });
}

// Appends the codes of one chunk to the staging ring, MSB first (BitWriter, rice.rs:162-208).
// Branch-free in the common case: every sample shifts its code into a 64-bit accumulator and a
// predicated store drops a 32-bit word whenever one is complete.  Words that only this thread
// writes are stored plainly; its first word (shared with the preceding threads) is kept in a
// register and OR-ed in at the end together with the last, partial word.
template <bool WINDOWED, bool MASKED>
__device__ __forceinline__ void emit_chunk(u32 *ring, const u32 (&u)[CH], int nv, int k, bool raw, u64 start,
                                           u32 wlo, u32 whi) {
    const u32 w0 = (u32)(start >> 5);
    u32 w = w0;
    int nb = (int)(start & 31);                    // leading zero bits stand for the part of w0 that is not ours
    u64 acc = 0;
    u32 head = 0;
    auto put = [&](u32 v, int len) {               // 1 <= len <= 32, v < 2^len
        acc = (acc << len) | v;
        nb += len;
        const bool full = nb >= 32;
        const u32 word = (u32)(acc >> ((nb - 32) & 63));
        const bool mine = full && w != w0 && (!WINDOWED || (w >= wlo && w < whi));
        if (mine) ring[w & (RING_WORDS - 1)] = word;
        head = (full && w == w0) ? word : head;
        w += full ? 1u : 0u;
        nb -= full ? 32 : 0;
    };
    const u32 kmask = (1u << k) - 1u;
    // Taken branches stall instruction fetch, so the per-sample code is straight-line: raw and Rice have
    // separate loops, and quotients above 16 (codes longer than 32 bits) are detected once per chunk.
    if (raw) {
#pragma unroll
        for (int j = 0; j < CH; j++)
            if (!MASKED || j < nv) put(u[j], 16);
    } else {
        u32 qmax = 0;
#pragma unroll
        for (int j = 0; j < CH; j++) qmax = max(qmax, (!MASKED || j < nv) ? (u[j] >> k) : 0u);
        if (qmax <= 16) {
#pragma unroll
            for (int j = 0; j < CH; j++) {             // encode_sample, rice.rs:94-114
                if (!MASKED || j < nv) {
                    const u32 q = u[j] >> k;
                    put((((1u << q) - 1u) << (k + 1)) | (u[j] & kmask), (int)q + k + 1);
                }
            }
        } else {
#pragma unroll 1
            for (int j = 0; j < (MASKED ? nv : CH); j++) {
                u32 uj = u[0];
#pragma unroll
                for (int t = 1; t < CH; t++) uj = (t == j) ? u[t] : uj;
                u32 q = uj >> k;
                while (q > 16) { const u32 t = q < 24 ? q : 24; put((1u << t) - 1u, (int)t); q -= t; }
                put((((1u << q) - 1u) << (k + 1)) | (uj & kmask), (int)q + k + 1);
            }
        }
    }
    const u32 tailw = nb > 0 ? (u32)(acc << (32 - nb)) : 0u;
    const bool in0 = !WINDOWED || (w0 >= wlo && w0 < whi);
#ifdef NO_ATOMIC_TEST
    if (w == w0) {
        if (in0 && tailw) ring[w0 & (RING_WORDS - 1)] = tailw;
    } else {
        if (in0 && head) ring[w0 & (RING_WORDS - 1)] = head;
        if (tailw && (!WINDOWED || (w >= wlo && w < whi))) ring[w & (RING_WORDS - 1)] = tailw;
    }
#else
    if (w == w0) {
        if (in0 && tailw) atomicOr(&ring[w0 & (RING_WORDS - 1)], tailw);
    } else {
        if (in0 && head) atomicOr(&ring[w0 & (RING_WORDS - 1)], head);
        if (tailw && (!WINDOWED || (w >= wlo && w < whi))) atomicOr(&ring[w & (RING_WORDS - 1)], tailw);
    }
#endif
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
template <int P>
__device__ void pack_channel(Smem &s, const ChanState &cs, const ChanResult &cr, uint8_t *obase, u64 pos, u32 *err,
                             unsigned long long *phase) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = cs.n;
    const int mode = cr.kind == 0 ? 13 : cr.order;
    const bool raw = cr.kind == 0;
    const int k = cr.k;
    const u64 abase = pos & ~3ull;
    const u64 lo = pos, hi = pos + cr.nbytes;
    u64 bitpos = (pos & 3ull) * 8ull;
    u32 wfl = 0;
    const int per_sc = NT * CH;
    const int nsc = (n + per_sc - 1) / per_sc;
    const double *qd = s.wqd;
    for (int sc = 0; sc < nsc; sc++) {
        const long long pk0 = clock64();
        const int i0 = sc * per_sc + tid * CH;
        const bool last = sc == nsc - 1;               // only the last super-chunk has short or absent chunks
        const int nv = last ? max(0, min(CH, n - i0)) : CH;
        u32 u[CH];
        u32 tb = 0;
        if (nv > 0) {
            switch (mode) {
                case 0: chunk_codes<0>(cs, i0, qd, u); break;
                case 1: chunk_codes<1>(cs, i0, qd, u); break;
                case 2: chunk_codes<2>(cs, i0, qd, u); break;
                case 3: chunk_codes<3>(cs, i0, qd, u); break;
                case 4: chunk_codes<4>(cs, i0, qd, u); break;
                case 13: chunk_codes<13>(cs, i0, qd, u); break;
                default:
                    if constexpr (P > 0) {
                        ForOrders<5, P>::run([&](auto oc) {
                            constexpr int O = decltype(oc)::value;
                            if (mode == O) chunk_codes<O>(cs, i0, qd, u);
                        });
                    }
                    break;
            }
            if (raw) {
                tb = 16u * (u32)nv;
            } else if (!last) {
                u32 qs = 0;
#pragma unroll
                for (int j = 0; j < CH; j++) qs += u[j] >> k;
                tb = qs + (u32)CH * (1u + (u32)k);
            } else {
#pragma unroll
                for (int j = 0; j < CH; j++)
                    if (j < nv) tb += (u[j] >> k) + 1u + (u32)k;
            }
        }
        const long long pk1 = clock64();
        // block exclusive scan of the chunk bit counts: warp scan, one barrier, then every warp scans the
        // 16 warp totals itself.  The totals are double buffered by round parity, and the ring words flushed
        // in the previous round are only touched again behind this round's barrier.
        u32 inc = tb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        u32 *tot = s.scan_warp[sc & 1];
        if (lane == 31) tot[wid] = inc;
        __syncthreads();
        u32 wv = lane < NWARP ? tot[lane] : 0;
        u32 wincl = wv;
#pragma unroll
        for (int o = 1; o < NWARP; o <<= 1) {
            u32 t = __shfl_up_sync(0xffffffffu, wincl, o);
            if (lane >= o) wincl += t;
        }
        const u32 warp_excl = __shfl_sync(0xffffffffu, wincl - wv, wid);
        const u32 scan_total = __shfl_sync(0xffffffffu, wincl, NWARP - 1);
        const long long pk2 = clock64();
        const u64 start = bitpos + warp_excl + (inc - tb);
        const u64 end_sc = bitpos + scan_total;
        const u32 wlast = (u32)((end_sc + 31) >> 5);
        u32 wlo = wfl;
        if (wlast - wlo <= (u32)RING_WORDS) {
            // common case: the whole super-chunk fits the staging ring
            if (!last) emit_chunk<false, false>(s.ring, u, CH, k, raw, start, 0, 0);
            else if (nv > 0) emit_chunk<false, true>(s.ring, u, nv, k, raw, start, 0, 0);
            __syncthreads();
            if (tid == 0) { const long long pk3 = clock64(); atomicAdd(phase + 5, (u64)(pk1 - pk0)); atomicAdd(phase + 6, (u64)(pk2 - pk1)); atomicAdd(phase + 7, (u64)(pk3 - pk2)); }
            const u32 wend = (u32)(end_sc >> 5);
            const long long pk4 = clock64();
            flush_ring(s.ring, obase, abase, lo, hi, wlo, wend);
            if (tid == 0) atomicAdd(phase + 14, (u64)(clock64() - pk4));
            wlo = wend;
        } else {
            for (;;) {
                const u32 whi = wlo + RING_WORDS;
                if (nv > 0) emit_chunk<true, true>(s.ring, u, nv, k, raw, start, wlo, whi);
                __syncthreads();
                const u32 wend = min(whi, (u32)(end_sc >> 5));
                flush_ring(s.ring, obase, abase, lo, hi, wlo, wend);
                __syncthreads();
                wlo = wend;
                if (whi >= wlast) break;
            }
        }
        wfl = wlo;
        bitpos = end_sc;
    }
    __syncthreads();
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
// ingest: quantise, silence test, deinterleave into 16-bit planes, mid/side energies
// ----------------------------------------------------------------------------
// channel header bytes inside an ALPC frame, writer.rs:272-299 (none in a Raw-typed frame, :267-270)
__device__ __forceinline__ u32 chan_hdr_bytes(bool all_raw, const ChanResult &r) {
    if (all_raw) return 0;
    return 1u + 4u * (r.kind == 2 ? (u32)r.order : 0u) + 1u + 1u + (r.kind != 0 ? 1u : 0u);
}
__device__ __forceinline__ void put_u32le(uint8_t *p, u32 v) {
    p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24);
}
__device__ __forceinline__ u32 pack2(i32 a, i32 b) { return ((u32)a & 0xffffu) | ((u32)b << 16); }

struct IngestAcc {
    bool loud = false;
    i64 vl = 0, vr = 0, vs = 0;
    __device__ __forceinline__ void pair(float a, float b, i32 &l, i32 &r) {
        loud |= is_loud(a) | is_loud(b);
        l = f32_to_i32(a); r = f32_to_i32(b);
        const i32 sd = l - r;
        vl += (i64)l * l; vr += (i64)r * r; vs += (i64)sd * sd;     // encoder.rs:136-149
    }
};

template <typename T>
__device__ void ingest_frame(Smem &s, const T *in, u32 len, u32 C, int16_t *planes, u32 stride) {
    const int tid = threadIdx.x;
    IngestAcc A;
    if (C == 2) {
        const u32 nf = len >> 1;
        u32 done = 0;
        // vector path: 8 sample-frames per thread step, 16-byte loads, 16-byte plane stores
        if ((reinterpret_cast<uintptr_t>(in) & 15) == 0) {
            const u32 ngrp = nf >> 3;
            auto convert = [&](const float (&f)[16], u32 gI) {
                i32 l[8], r[8];
#pragma unroll
                for (int i = 0; i < 8; i++) A.pair(f[2 * i], f[2 * i + 1], l[i], r[i]);
                *reinterpret_cast<uint4 *>(planes + (size_t)gI * 8) =
                    make_uint4(pack2(l[0], l[1]), pack2(l[2], l[3]), pack2(l[4], l[5]), pack2(l[6], l[7]));
                *reinterpret_cast<uint4 *>(planes + stride + (size_t)gI * 8) =
                    make_uint4(pack2(r[0], r[1]), pack2(r[2], r[3]), pack2(r[4], r[5]), pack2(r[6], r[7]));
            };
            // two groups per step: both groups' 16-byte loads are issued before either is converted
            for (u32 gI = tid; gI < ngrp; gI += 2 * NT) {
                const u32 gJ = gI + NT;
                const bool two = gJ < ngrp;
                float f[16], h[16];
                if constexpr (sizeof(T) == 4) {
                    const float4 *p = reinterpret_cast<const float4 *>(in) + (size_t)gI * 4;
                    const float4 *q = reinterpret_cast<const float4 *>(in) + (size_t)(two ? gJ : gI) * 4;
                    const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
                    const float4 a2 = __ldg(q), b2 = __ldg(q + 1), c2 = __ldg(q + 2), d2 = __ldg(q + 3);
                    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
                    f[8] = c.x; f[9] = c.y; f[10] = c.z; f[11] = c.w; f[12] = d.x; f[13] = d.y; f[14] = d.z; f[15] = d.w;
                    h[0] = a2.x; h[1] = a2.y; h[2] = a2.z; h[3] = a2.w; h[4] = b2.x; h[5] = b2.y; h[6] = b2.z; h[7] = b2.w;
                    h[8] = c2.x; h[9] = c2.y; h[10] = c2.z; h[11] = c2.w; h[12] = d2.x; h[13] = d2.y; h[14] = d2.z; h[15] = d2.w;
                } else {
                    const int4 *p = reinterpret_cast<const int4 *>(in) + (size_t)gI * 2;
                    const int4 *q = reinterpret_cast<const int4 *>(in) + (size_t)(two ? gJ : gI) * 2;
                    const int4 a = __ldg(p), b = __ldg(p + 1), a2 = __ldg(q), b2 = __ldg(q + 1);
                    const int wv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                    const int wu[8] = {a2.x, a2.y, a2.z, a2.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        f[2 * i] = pcm_to_f32((int)(int16_t)(wv[i] & 0xffff));
                        f[2 * i + 1] = pcm_to_f32(wv[i] >> 16);
                        h[2 * i] = pcm_to_f32((int)(int16_t)(wu[i] & 0xffff));
                        h[2 * i + 1] = pcm_to_f32(wu[i] >> 16);
                    }
                }
                convert(f, gI);
                if (two) convert(h, gJ);
            }
            done = ngrp << 3;
        }
        for (u32 i = done + tid; i < nf; i += NT) {
            i32 l, r;
            A.pair(sample_f32<T>(in, 2 * (size_t)i), sample_f32<T>(in, 2 * (size_t)i + 1), l, r);
            planes[i] = (int16_t)l;
            planes[stride + i] = (int16_t)r;
        }
        if ((len & 1) && tid == 0) {           // ragged tail: channel 0 gets one more sample (encoder.rs:84-90)
            const float a = sample_f32<T>(in, (size_t)len - 1);
            A.loud |= is_loud(a);
            planes[nf] = (int16_t)f32_to_i32(a);
        }
    } else if (C == 1) {
        u32 done = 0;
        if ((reinterpret_cast<uintptr_t>(in) & 15) == 0) {
            const u32 ngrp = len >> 3;
            for (u32 gI = tid; gI < ngrp; gI += NT) {
                float f[8];
                if constexpr (sizeof(T) == 4) {
                    const float4 *p = reinterpret_cast<const float4 *>(in) + (size_t)gI * 2;
                    const float4 a = __ldg(p), b = __ldg(p + 1);
                    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
                } else {
                    const int4 a = __ldg(reinterpret_cast<const int4 *>(in) + gI);
                    const int wv[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        f[2 * i] = pcm_to_f32((int)(int16_t)(wv[i] & 0xffff));
                        f[2 * i + 1] = pcm_to_f32(wv[i] >> 16);
                    }
                }
                i32 v[8];
#pragma unroll
                for (int i = 0; i < 8; i++) { A.loud |= is_loud(f[i]); v[i] = f32_to_i32(f[i]); }
                *reinterpret_cast<uint4 *>(planes + (size_t)gI * 8) =
                    make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
            }
            done = ngrp << 3;
        }
        for (u32 i = done + tid; i < len; i += NT) {
            const float a = sample_f32<T>(in, i);
            A.loud |= is_loud(a);
            planes[i] = (int16_t)f32_to_i32(a);
        }
    } else {
        for (u32 e = tid; e < len; e += NT) {
            const float a = sample_f32<T>(in, e);
            A.loud |= is_loud(a);
            const u32 c = e % C, i = e / C;
            planes[(size_t)c * stride + i] = (int16_t)f32_to_i32(a);
        }
    }
    // zero the padding behind each channel (chunk loads read up to the next multiple of 16)
    for (u32 c = 0; c < C; c++) {
        const u32 cl = len > c ? (len - c + C - 1) / C : 0;
        for (u32 i = cl + tid; i < stride; i += NT) planes[(size_t)c * stride + i] = 0;
    }
    // block-wide: loud flag and the three energies
    const int lane = tid & 31;
    const u32 anyloud = __ballot_sync(0xffffffffu, A.loud);
    if (C == 2) {
        const u64 a = warp_sum64((u64)A.vl), b = warp_sum64((u64)A.vr), c = warp_sum64((u64)A.vs);
        if (lane == 0) { atomic_add64(&s.ms_var[0], a); atomic_add64(&s.ms_var[1], b); atomic_add64(&s.ms_var[2], c); }
    }
    if (lane == 0 && anyloud) atomicOr(reinterpret_cast<u32 *>(&s.loud), 1u);
}

// ----------------------------------------------------------------------------
// the frame-encode kernel
// ----------------------------------------------------------------------------
extern __shared__ __align__(16) unsigned char dyn_smem[];

template <int P>
__global__ void __launch_bounds__(NT, FLO_VARIANT_CTAS) k_encode_frames(const EncodeParams p) {
    Smem &s = *reinterpret_cast<Smem *>(dyn_smem);
    int16_t *smem_planes = reinterpret_cast<int16_t *>(dyn_smem + ((sizeof(Smem) + 15) & ~size_t(15)));
    const int tid = threadIdx.x;
    for (int i = tid; i < RING_WORDS; i += NT) s.ring[i] = 0;
    if (tid < 8) s.cnt[tid] = 0;
    ChanResult *cres = p.cres + (size_t)blockIdx.x * 256;
    const int level = p.level;
    // P = LPC max order analysed by this instantiation (0: levels 0-3, fixed predictors only)
    const int PL = order_of_level(level);            // encoder.rs:289-302
    const int fmax = PL < 4 ? PL : 4;
    const bool lpc_on = P > 0;                        // level >= 3 && max_order > 4, encoder.rs:204
    const bool prune = p.report == nullptr;      // the parity report wants every candidate's exact size

    for (;;) {
        __syncthreads();
        if (tid == 0) {
            // A ticket is taken only when the CTA is ready to start the frame: frames are then published
            // (look-back status) in nearly ticket order.  Taking tickets one frame ahead (to prefetch the next
            // frame's samples into L2) was tried: a CTA that lags then holds an early ticket for a whole frame
            // time and every later frame stalls in the look-back -- 1.4 ms -> 2.0 ms per 1184 frames.
            s.g = p.frame_begin + atomicAdd(p.ticket, 1u);
            s.loud = 0; s.ms = 0;
            s.ms_var[0] = s.ms_var[1] = s.ms_var[2] = 0;
        }
        __syncthreads();
        const u32 g = s.g;
        if (g >= p.frame_end) break;
        const long long tc0 = clock64();
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

        const long long tcA = clock64();
        if (p.format == FLO_FMT_PCM16)
            ingest_frame<int16_t>(s, reinterpret_cast<const int16_t *>(tr.samples) + start, len, C, planes, stride);
        else
            ingest_frame<float>(s, reinterpret_cast<const float *>(tr.samples) + start, len, C, planes, stride);
        const long long tcB = clock64();
        __syncthreads();
        if (tid == 0) { atomicAdd(p.phase_cycles + 12, (u64)(tcA - tc0)); atomicAdd(p.phase_cycles + 13, (u64)(tcB - tcA)); }

        const u64 data_base = tr.static_off + FILE_HDR + 4ull + 20ull * tr.n_frames;   // writer.rs:51, 89-95

        if (!s.loud) {
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
            continue;
        }

        const long long tc1 = clock64();
        if (tid == 0) atomicAdd(&s.cnt[0], 1u);
        // mid/side decision, encoder.rs:94-100, 131-153
        int ms = 0;
        if (C == 2) {
            const i64 vl = (i64)s.ms_var[0], vr = (i64)s.ms_var[1], vs = (i64)s.ms_var[2];
            ms = vs < (vl + vr) / 2 ? 1 : 0;
            if (ms && (len & 1)) {             // the unpaired tail sample of L is dropped by the zip (encoder.rs:160)
                if (tid == 0) planes[len >> 1] = 0;
            }
        }

        // per-channel predictor search, GROUP channels at a time
        for (u32 c0 = 0; c0 < C; c0 += GROUP) {
            const int nch = (int)min((u32)GROUP, C - c0);
            __syncthreads();
            if (tid < nch) {
                ChanState &cs = s.cs[tid];
                const u32 c = c0 + tid;
                u32 cl = len > c ? (len - c + C - 1) / C : 0;
                if (ms) cl = len >> 1;                                      // zip in to_mid_side truncates, encoder.rs:160-167
                cs.n = (int)cl;
                cs.nfull = (int)(cl / CH);
                cs.tail = (int)(cl % CH);
                cs.msmode = ms ? (c == 0 ? 1 : 2) : 0;
                cs.pa = ms ? planes : planes + (size_t)c * stride;
                cs.pb = planes + stride;
                for (int o = 0; o < 5; o++) { cs.fix_sum[o] = 0; cs.fix_or[o] = 0; }
                for (int l = 0; l <= MAXORD; l++) cs.ac[l] = 0;
                for (int o = 0; o < NLPC; o++) { cs.l_sum[o] = 0; cs.l_or[o] = 0; cs.l_t0[o] = 0; cs.l_t1[o] = 0; cs.lpc_ok[o] = 0; }
                cs.ex_cand = -1; cs.ex_s = 0; cs.ex_max = 0; cs.ex_fixed = 0;
            }
            __syncthreads();
            const long long ta0 = clock64();
            bool any_lpc = false;
            for (int q = 0; q < nch; q++) any_lpc |= lpc_on && s.cs[q].n > 5;
            if (any_lpc) pass1<P>(s, nch);
            else pass1<0>(s, nch);
            __syncthreads();
            const long long ta1 = clock64();
            if ((tid >> 5) < nch && s.cs[tid >> 5].n > 0) after_pass1_warp<P>(s.cs[tid >> 5], fmax, lpc_on);
            __syncthreads();
            const long long ta2 = clock64();
            bool run2 = false;
            for (int q = 0; q < nch; q++)
                for (int o = 0; o < NLPC; o++) run2 |= s.cs[q].n > 0 && s.cs[q].lpc_ok[o] != 0;
            if (run2) {
                if constexpr (P > 0) pass2<P>(s, nch);
                __syncthreads();
            }
            const long long ta3 = clock64();
            if ((tid >> 5) < nch && s.cs[tid >> 5].n > 0) {
                ChanState &cs = s.cs[tid >> 5];
                if (run2) after_pass2_warp(cs, P, s.cnt);
                next_open_candidate_warp(cs, prune, s.cnt);
            }
            __syncthreads();
            // exact evaluation of whatever is still open (bounded candidates that can still win)
            for (;;) {
                bool more = false;
                for (int q = 0; q < nch; q++) more |= s.cs[q].n > 0 && s.cs[q].ex_cand >= 0;
                if (!more) break;
                if (tid == 0) atomicAdd(&s.cnt[1], 1u);
                pass3<P>(s, nch);
                __syncthreads();
                if ((tid >> 5) < nch && s.cs[tid >> 5].n > 0) {
                    ChanState &cs = s.cs[tid >> 5];
                    if ((tid & 31) == 0) after_pass3(cs);
                    __syncwarp();
                    next_open_candidate_warp(cs, prune, s.cnt);
                }
                __syncthreads();
            }
            if (tid == 0) {
                const long long ta4 = clock64();
                atomicAdd(p.phase_cycles + 8, (u64)(ta1 - ta0)); atomicAdd(p.phase_cycles + 9, (u64)(ta2 - ta1));
                atomicAdd(p.phase_cycles + 10, (u64)(ta3 - ta2)); atomicAdd(p.phase_cycles + 11, (u64)(ta4 - ta3));
            }
            // encode_channel_int, encoder.rs:184-216: strictly smaller wins, candidates in order
            if (tid < nch) {
                ChanState &cs = s.cs[tid];
                const u32 c = c0 + tid;
                ChanResult r;
                if (cs.n == 0) {
                    r.kind = 3; r.order = 0; r.k = 0; r.nbytes = 0; r.shift = 0;
                    for (int j = 0; j < MAXORD; j++) r.coef[j] = 0;
                } else {
                    i64 best = cs.cand_size[0];
                    int bj = 0;
                    for (int j = 1; j < NCAND; j++) {
                        if (cs.cand_state[j] != CS_EXACT) continue;
                        const i64 sz = cs.cand_size[j];
                        if (sz < best) { best = sz; bj = j; }
                    }
                    r.kind = bj == 0 ? 0 : (bj <= 5 ? 1 : 2);
                    r.order = bj == 0 ? 0 : bj - 1;
                    r.k = cs.cand_k[bj];
                    r.nbytes = (u32)best;
                    for (int j = 0; j < MAXORD; j++) r.coef[j] = (r.kind == 2 && j < r.order) ? cs.qc[r.order - 5][j] : 0;
                    r.shift = r.kind == 2 ? cs.lpc_shift[r.order - 5] : 0;
                }
                r.pad[0] = r.pad[1] = r.pad[2] = 0;
                cres[c] = r;
                if (p.report && c < REPORT_CH) {
                    flo_cand_report *rep = p.report + ((size_t)g * REPORT_CH + c) * NCAND;
                    for (int j = 0; j < NCAND; j++) {
                        const bool ex = cs.n > 0 && cs.cand_state[j] == CS_EXACT;
                        rep[j].k = ex ? cs.cand_k[j] : 0; rep[j].pad = 0; rep[j].size = ex ? cs.cand_size[j] : -1;
                    }
                }
            }
        }
        __syncthreads();

        const long long tc2 = clock64();
        // frame typing and size, encoder.rs:102-127, types.rs:242-267
        bool all_raw = true;
        u32 fsize = 6;
        const u32 frame_type_alpc = (PL >= 1 && PL <= 12) ? (u32)PL : 8u;     // FrameType::from_order, types.rs:69-85
        for (u32 c = 0; c < C; c++)
            if (cres[c].order > 0) all_raw = false;
        for (u32 c = 0; c < C; c++) {
            const ChanResult &r = cres[c];
            fsize += 4 + chan_hdr_bytes(all_raw, r) + r.nbytes;
        }
        if (tid < 32) {
            u64 ex = lookback_exclusive(p.status, g, fsize);
            if (tid == 0) { s.frame_excl = ex; p.frame_excl[g] = ex; p.frame_size[g] = fsize; }
        }
        __syncthreads();

        const long long tc3 = clock64();
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
            const u32 hdr = chan_hdr_bytes(all_raw, r);
            __syncthreads();
            if (tid == 0) {
                ChanState &cs = s.cs[0];
                u32 cl = len > c ? (len - c + C - 1) / C : 0;
                if (ms) cl = len >> 1;
                cs.n = (int)cl;
                cs.msmode = ms ? (c == 0 ? 1 : 2) : 0;
                cs.pa = ms ? planes : planes + (size_t)c * stride;
                cs.pb = planes + stride;
            }
            if (tid < MAXORD) { s.wcoef[tid] = r.coef[tid]; s.wqd[tid] = ldexp((double)r.coef[tid], -r.shift); }
            __syncthreads();
            if (tid == 0) {
                put_u32le(o + pos, hdr + r.nbytes);
                if (!all_raw) {
                    uint8_t *h = o + pos + 4;
                    const u32 nco = r.kind == 2 ? (u32)r.order : 0u;
                    *h++ = (uint8_t)nco;
                    for (u32 j = 0; j < nco; j++) { put_u32le(h, (u32)r.coef[j]); h += 4; }
                    *h++ = (uint8_t)(r.kind == 2 ? r.shift : (r.kind == 1 ? 128 + r.order : 0));   // encoder.rs:243, 279
                    *h++ = (uint8_t)(r.kind == 0 ? 2 : 0);                                         // ResidualEncoding
                    if (r.kind != 0) *h++ = (uint8_t)r.k;
                }
            }
            if (r.kind != 3 && r.nbytes > 0) pack_channel<P>(s, s.cs[0], r, o, pos + 4 + hdr, p.err, p.phase_cycles);
            pos += 4 + hdr + r.nbytes;
        }
        if (tid == 0 && pos - fpos != fsize) atomicExch(p.err, 0xBAD00002u);
        if (tid == 0) {
            const long long tc4 = clock64();
            atomicAdd(p.phase_cycles + 0, (u64)(tc1 - tc0)); atomicAdd(p.phase_cycles + 1, (u64)(tc2 - tc1));
            atomicAdd(p.phase_cycles + 2, (u64)(tc3 - tc2)); atomicAdd(p.phase_cycles + 3, (u64)(tc4 - tc3));
            atomicAdd(p.phase_cycles + 4, (u64)(tc4 - tc0));
        }
    }
    __syncthreads();
    if (tid < 8 && s.cnt[tid]) atomicAdd(p.counters + tid, s.cnt[tid]);
}

// ----------------------------------------------------------------------------
// launch glue of this variant
// ----------------------------------------------------------------------------
static cudaError_t variant_configure(size_t dyn_smem) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_encode_frames<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem))) return e;
    if ((e = cudaFuncSetAttribute(k_encode_frames<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem))) return e;
    if ((e = cudaFuncSetAttribute(k_encode_frames<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem))) return e;
    if ((e = cudaFuncSetAttribute(k_encode_frames<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem))) return e;
    return cudaFuncSetAttribute(k_encode_frames<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
}
static cudaError_t variant_launch(const EncodeParams &p, int grid, size_t dyn_smem, cudaStream_t st) {
    if (p.frame_end <= p.frame_begin) return cudaSuccess;
    // one instantiation per LPC max order (encoder.rs:289-302); levels 0-3 never try LPC (encoder.rs:204)
    switch (p.level) {
        case 4: k_encode_frames<6><<<grid, NT, dyn_smem, st>>>(p); break;
        case 5: case 6: k_encode_frames<8><<<grid, NT, dyn_smem, st>>>(p); break;
        case 7: k_encode_frames<10><<<grid, NT, dyn_smem, st>>>(p); break;
        case 8: case 9: k_encode_frames<12><<<grid, NT, dyn_smem, st>>>(p); break;
        default: k_encode_frames<0><<<grid, NT, dyn_smem, st>>>(p); break;
    }
    return cudaGetLastError();
}
static int variant_occupancy(size_t dyn_smem) {
    int n = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_encode_frames<8>, NT, dyn_smem);
    return n;
}
