// This file is the body of one build variant of the frame-encode kernel: it is included by
// flo_encode_nt{512,256,128}.cu inside namespace flo::FLO_VARIANT_NS with FLO_VARIANT_NT threads per CTA.
//
// Round-2 formulation.  What changed against the round-1 body (profiles/r01_*):
//   * LPC FIR: the predictions of a block of samples (and of two LPC orders at a time) are accumulated as
//     independent FP64 chains that are interleaved in the instruction stream -- round 1 issued each
//     sample's chain of O dependent DFMAs back to back and waited for every one of them;
//   * the two orders of a sweep share one converted sample window (fewer int -> f64 conversions);
//   * the first 16 samples of a channel (warm-up samples, lpc.rs:283-285 / 311-352) and the < 16 samples
//     behind the last full chunk are analysed by single lanes through the scalar reference arithmetic, so the
//     unrolled bodies carry neither warm-up selects nor history tests;
//   * ingest: the interleaved input of a frame is staged through shared memory by 1-D bulk async copies
//     (cp.async.bulk + mbarrier, up to four stages in flight) issued by one elected thread, and the frame a
//     CTA is likely to take next is prefetched into L2 while the current one is packed;
//   * bit packer: straight-line append with an address-swap for the shared first word (no per-code head
//     test), the staging ring is re-based per round and leaves the CTA as 16-byte stores;
//   * fixed-predictor statistics only for the orders the level can choose (encoder.rs:193: 0..=min(4, P));
//   * profiling clocks are compiled in only with -DFLO_PHASE_CLOCKS (tools/phases.py).
constexpr int NT = FLO_VARIANT_NT;
constexpr int NWARP = NT / 32;

#ifndef FLO_SINGLE_BS
#define FLO_SINGLE_BS 4       // samples per block of a one-order FIR sweep (independent chains)
#endif
#ifndef FLO_PACK_BS
#define FLO_PACK_BS 4         // independent FIR chains of the packer's (and pass 3's) residual recomputation (8: no change, 3.06 ms)
#endif
#ifndef FLO_PAIR_SWEEPS
#define FLO_PAIR_SWEEPS 0     // 1: two LPC orders per sweep share one window (fewer loads, but the second set of chains and
#endif                        //    coefficients spills ~1.5 KB per thread at 128 registers: 3.13 vs 3.08 ms at level 5)
#ifndef FLO_PAIR_BS
#define FLO_PAIR_BS 4         // samples per block of the two-order FIR sweeps (independent chains = 2 x this)
#endif
#ifdef FLO_PASSES_NOINLINE    // experiment: the passes as real calls (their own register allocation)
#define PASS_FN __device__ __noinline__
#else
#define PASS_FN __device__
#endif
#ifdef FLO_PHASE_CLOCKS
#define PH(...) __VA_ARGS__
#else
#define PH(...)
#endif
#ifdef FLO_LEV_CLOCKS         // experiment: Levinson sub-steps on the pack slots 5..7 and 14
#define LV(...) __VA_ARGS__
#define LEVC 1
__device__ unsigned long long *g_lev_phase;
#else
#define LV(...)
#define LEVC 0
#endif

// ----------------------------------------------------------------------------
// shared state of the frame-encode CTA
// ----------------------------------------------------------------------------
constexpr int GROUP = 2;              // channels analysed jointly (stereo = one group)
constexpr int NLPC = MAXORD - 4;      // LPC orders 5..12
constexpr int MAX_STAGES = 4;         // bulk-copy stages of the ingest ring
constexpr int SPT = NT >= 512 ? 4 : 8;   // sample frames per thread per ingest step
constexpr int WRING = 512;            // words of one warp's staging ring in the packer

// candidate states
constexpr int CS_ABSENT = 0, CS_EXACT = 1, CS_BOUNDED = 2;

struct ChanState {
    // layout of the coded channel (after the mid/side decision)
    const int16_t *pa, *pb;
    int msmode;                       // 0 plain, 1 mid = L + R, 2 side = L - R (encoder.rs:156-170)
    int sel_lo, sel_hi;               // dp2a byte selectors that add (mid) or subtract (side) plane b
    int glob;                         // planes live in global memory (L2-resident scratch), not in shared memory
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
    i32 cand_src[NCAND];              // where the exact size came from: 0 pass 3 (partS), 1 / 2 pass 2 window t0 / t1 (partT)
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
    double wqd[GROUP][MAXORD];
    u32 scan_warp[2][NWARP];
    u32 cnt[8];                       // analysis counters of this CTA (flushed at kernel end)
    u32 g;                            // current global frame
    i32 ms;                           // mid/side chosen (encoder.rs:94-100)
    i32 loud;
    u64 ms_var[3];
    u64 frame_excl;                   // exclusive prefix of frame sizes
    unsigned long long bar_full[MAX_STAGES];   // mbarriers of the ingest stages
    u32 headw[NT];                    // packer: first (shared) word of each thread's chunk stream
    // per-region (= per warp of the channel) partial sums of the analysis passes: the packer's regions start at
    // bit offsets that follow from them, so the warps of a channel pack independently of each other
    u64 partA[GROUP][NCAND][NWARP];   // sum|r| per region (fixed: pass 1, LPC: pass 2)
    u64 partT[GROUP][NLPC][2][NWARP]; // pass 2: sum(w >> j0), sum(w >> (j0 + 1)) per region
    u64 partS[GROUP][NCAND][NWARP];   // pass 3: exact sum(w >> j) per region
    u64 partAC[GROUP][MAXORD + 1][NWARP];   // pass 1: autocorrelation lags per region
    // (a region's slots are written by its own warp only -- plain read-modify-write, no 64-bit shared-memory
    //  atomics, which are compare-and-swap loops; the channel totals are summed over the regions afterwards)
    u32 edge[GROUP][NWARP + 1];       // packer: bits of the words shared by two regions, by boundary
    i32 redo[GROUP];                  // a channel asks for the second run of pass 2
    ChanResult wres[GROUP];           // the winners of the channels being packed (copied from global memory once)
    // CRC32 of the frame's bytes (crc32.rs): tables staged from global memory once per CTA
    u32 crc_x[4][256];                // v -> v * x^(32 NT) mod p, by byte of v (the strided Horner step)
    u32 crc_klane[NT];                // x^(32 (NT - t)) mod p
    u32 crc_slice0[256];              // the byte-wise table of crc32.rs:2-20
    u32 fcrc;                         // raw CRC state of the current frame, XORed together by the warps
};
// dynamic shared memory: Smem | work area (ingest stages, then the packer's staging ring) | sample planes
constexpr size_t SMEM_HDR = (sizeof(Smem) + 127) & ~size_t(127);

static size_t encode_static_smem() { return SMEM_HDR; }

__device__ __forceinline__ u64 warp_sum64(u64 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void atomic_add64(u64 *p, u64 v) { atomicAdd(reinterpret_cast<unsigned long long *>(p), v); }
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_shared_u32(u32 addr, u32 v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ u32 ld_shared_u32(u32 addr) { u32 v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory"); return v; }
__device__ __forceinline__ void atom_or_shared(u32 addr, u32 v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// ---- bulk async copy (TMA, 1-D) + mbarrier ----------------------------------------------
__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(u32 dst, const void *src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void *p, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

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
    LV(const long long l0 = clock64();)
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
    LV(const long long l1 = clock64();)
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
    LV(const long long l2 = clock64();)
    // (3) quantise every (order, j) pair in parallel (lpc.rs:263-273); taps of orders that were not reached are
    //     zero so that the FIR sweeps can run them blindly (their statistics are ignored)
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
        } else {
            cs.qc[t][j] = 0;
            cs.qd[t][j] = 0.0;
        }
    }
    __syncwarp();
    LV(if (threadIdx.x == 0) { const long long l3 = clock64(); atomicAdd(g_lev_phase + 5, (u64)(l1 - l0)); atomicAdd(g_lev_phase + 6, (u64)(l2 - l1)); atomicAdd(g_lev_phase + 7, (u64)(l3 - l2)); })
}

// ----------------------------------------------------------------------------
// sample access: 16 samples of the coded channel starting at i0 (multiple of 16) plus NH samples of
// history.  x[NH + j] = s[i0 + j], x[NH - 1 - h] = s[i0 - 1 - h].  HIST = false: the chunk is the first of
// the frame and its history is zero.  Planes are zero padded behind the channel end.  16-bit pairs become
// two i32 through the integer dot-product unit: dp2a_lo(w, b, c) = c + w.lo * b.byte0 + w.hi * b.byte1, so
// b = 0x0001 picks the low half, 0x0100 the high half and 0x00ff / 0xff00 subtract them (side = L - R).
// ----------------------------------------------------------------------------
template <int NH, bool HIST>
__device__ __forceinline__ void load_chunk(const ChanState &cs, int i0, i32 (&x)[NH + CH]) {
    static_assert(CH == 16, "chunk of 16 samples");
    static_assert(NH == 0 || NH == 4 || NH == 8 || NH == 12, "history of 0, 4, 8 or 12 samples");
    constexpr int HW = NH / 2;                         // history words (sample pairs)
    // All loads of both planes are issued before the first use: with the planes in the L2-resident scratch a
    // load costs a full L2 round trip, and one round trip per chunk is what the other warps can cover.
    const int4 z = make_int4(0, 0, 0, 0);
    const bool ms = cs.msmode != 0;
    const int4 *pa = reinterpret_cast<const int4 *>(cs.pa + i0);
    const int4 *pb = reinterpret_cast<const int4 *>(cs.pb + i0);
    int4 a0 = pa[0], a1 = pa[1], a2 = z, a3 = z, b0 = z, b1 = z, b2 = z, b3 = z;
    if constexpr (HIST && HW > 0) a2 = pa[-1];
    if constexpr (HIST && HW > 4) a3 = pa[-2];
    if (ms) {
        b0 = pb[0]; b1 = pb[1];
        if constexpr (HIST && HW > 0) b2 = pb[-1];
        if constexpr (HIST && HW > 4) b3 = pb[-2];
    }
    const int wa[16] = {a3.x, a3.y, a3.z, a3.w, a2.x, a2.y, a2.z, a2.w, a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int i = 0; i < HW + 8; i++) {
        x[2 * i] = __dp2a_lo(wa[8 - HW + i], 0x0001, 0);
        x[2 * i + 1] = __dp2a_lo(wa[8 - HW + i], 0x0100, 0);
    }
    if (ms) {                                          // mid = L + R, side = L - R (encoder.rs:156-170)
        const int slo = cs.sel_lo, shi = cs.sel_hi;
        const int wb[16] = {b3.x, b3.y, b3.z, b3.w, b2.x, b2.y, b2.z, b2.w, b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < HW + 8; i++) {
            x[2 * i] = __dp2a_lo(wb[8 - HW + i], slo, x[2 * i]);
            x[2 * i + 1] = __dp2a_lo(wb[8 - HW + i], shi, x[2 * i + 1]);
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

// fixed_predictor_residuals, lpc.rs:301-359, orders 0..NF-1 over one chunk: r_o[i] = o-th difference.
// FIRST (first chunk of the frame, zero history): r_o[i] = i-th difference for i < o (lpc.rs:311-352).
// x: 4 history samples + CH (no history when NF == 1).  fn(j, d[0..NF-1]).
template <int NF, bool FIRST, class F>
__device__ __forceinline__ void fixed_chunk(const i32 *x, F &&fn) {
    constexpr int H = NF > 1 ? 4 : 0;
    i32 xp = 0, p1 = 0, p2 = 0, p3 = 0;
    if constexpr (NF > 1) {
        xp = x[3];
        p1 = x[3] - x[2];
        const i32 p1b = x[2] - x[1], p1c = x[1] - x[0];
        p2 = p1 - p1b;
        const i32 p2b = p1b - p1c;
        p3 = p2 - p2b;
    }
#pragma unroll
    for (int j = 0; j < CH; j++) {
        i32 d[5];
        d[0] = x[H + j];
        d[1] = d[0] - xp;
        d[2] = d[1] - p1;
        d[3] = d[2] - p2;
        d[4] = d[3] - p3;
        xp = d[0]; p1 = d[1]; p2 = d[2]; p3 = d[3];
        if constexpr (FIRST) {
            if (j < 4) {
#pragma unroll
                for (int o = 4; o > j; o--) d[o] = d[j];
            }
        }
        fn(j, d);
    }
}

// int -> f64 on the conversion pipe.  `volatile` pins each conversion where it is written, so the
// compiler neither keeps a whole chunk of doubles alive (spills) nor re-converts a sample per use.
__device__ __forceinline__ double cvt_f64(i32 x) {
#ifdef FLO_CVT_MAGIC
    // experiment: 2^52 + 2^31 + x built from its bit pattern, minus the constant -- one ALU and one FP64-pipe
    // instruction instead of I2F.F64 on the XU pipe; exact for every i32.  Measured slower (3.20 vs 3.07 ms at level 5):
    // the XU pipe (22 % busy) is not what the kernel waits for, the extra issue slots are.  Off.
    return __dadd_rn(__hiloint2double(0x43300000, (int)((u32)x ^ 0x80000000u)), -4503601774854144.0);
#else
    double d;
    asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d) : "r"(x));
    return d;
#endif
}

// calc_residuals_int, lpc.rs:279-298, on the FP64 pipe, for LPC orders OA and OB (OB = 0: one order) over
// one chunk.  With c[t] = q[t] / 2^shift (exact) a chain of fused multiply-adds holds
// sum(q[t] * s[i-1-t]) / 2^shift exactly (|sum q s| < 2^53), so floor() of it equals the reference's
// arithmetic `pred >> shift`; adding 1.5 * 2^52 with round-down leaves that floor, modulo 2^32, in the low
// word -- the `pred as i32` truncation.  The predictions of BS consecutive samples (times two orders) are
// independent chains, written tap-major so that consecutive instructions never depend on each other; both
// orders read one window of converted samples.  x: NH >= OA history samples + CH.  fa(j, r), fb(j, r).
template <int NH, int OA, int OB, int BS, bool FIRST, class FA, class FB>
__device__ __forceinline__ void lpc_chunk2(const i32 (&x)[NH + CH], const double *qa, const double *qb, FA &&fa, FB &&fb) {
    static_assert(OA >= OB && OA <= NH && CH % BS == 0 && BS % 2 == 0, "orders / block");
    double ca[OA], cb[OB > 0 ? OB : 1];
#pragma unroll
    for (int t = 0; t < OA; t++) ca[t] = qa[t];
#pragma unroll
    for (int t = 0; t < OB; t++) cb[t] = qb[t];
    double w[OA + CH];                                 // w[i] = f64(x[NH - OA + i]); sample j sits at w[OA + j]
#pragma unroll
    for (int t = 0; t < OA + BS - 1; t++) w[t] = cvt_f64(x[NH - OA + t]);
#pragma unroll
    for (int b = 0; b < CH / BS; b++) {
        double pa[BS], pb[BS];
#pragma unroll
        for (int s = 0; s < BS; s++) { pa[s] = 0.0; pb[s] = 0.0; }
#pragma unroll
        for (int t = 0; t < OA; t++) {
#pragma unroll
            for (int s = 0; s < BS; s++) pa[s] = __fma_rn(ca[t], w[OA + b * BS + s - 1 - t], pa[s]);
            if (t < OB) {
#pragma unroll
                for (int s = 0; s < BS; s++) pb[s] = __fma_rn(cb[t], w[OA + b * BS + s - 1 - t], pb[s]);
            }
        }
        // conversions for the next block overlap this block's statistics
#pragma unroll
        for (int s = 0; s < BS; s++) {
            const int idx = b * BS + BS - 1 + s;
            if (idx < CH - 1) w[OA + idx] = cvt_f64(x[NH + idx]);
        }
#pragma unroll
        for (int s = 0; s < BS; s++) {
            const int j = b * BS + s;
            const i32 xj = x[NH + j];
            i32 ra = (i32)((u32)xj - (u32)__double2loint(__dadd_rd(pa[s], 6755399441055744.0)));
            if (FIRST && j < OA) ra = xj;              // warm-up, lpc.rs:283-285
            fa(j, ra);
            if constexpr (OB > 0) {
                i32 rb = (i32)((u32)xj - (u32)__double2loint(__dadd_rd(pb[s], 6755399441055744.0)));
                if (FIRST && j < OB) rb = xj;
                fb(j, rb);
            }
        }
    }
}

// ----------------------------------------------------------------------------
// scalar residuals at one sample index (the first chunk of a channel and the < 16 samples behind the
// last full chunk are analysed one sample per lane through these, in the reference's own arithmetic)
// ----------------------------------------------------------------------------
__device__ __forceinline__ i32 sample_at(const ChanState &cs, int i) {
    if (i < 0 || i >= cs.n) return 0;
    const i32 a = cs.pa[i];
    if (cs.msmode == 0) return a;
    const i32 b = cs.pb[i];
    return cs.msmode == 1 ? a + b : a - b;
}
// The lane's sample and its MAXORD predecessors, fetched with independent loads (one L2 round trip instead
// of one per tap): w[t] = s[i - t], zero outside the channel.
struct LaneWin { i32 w[MAXORD + 1]; };
__device__ __forceinline__ void lane_window(const ChanState &cs, int i, LaneWin &win) {
#pragma unroll
    for (int t = 0; t <= MAXORD; t++) win.w[t] = sample_at(cs, i - t);
}
// fixed_predictor_residuals (lpc.rs:301-359) at index i: the min(o, i)-th finite difference
__device__ __forceinline__ i32 fixed_residual_at(const LaneWin &win, int o, int i) {
    const u32 w0 = (u32)win.w[0], w1 = (u32)win.w[1], w2 = (u32)win.w[2], w3 = (u32)win.w[3], w4 = (u32)win.w[4];
    const int oo = o < i ? o : i;
    const u32 d1 = w0 - w1, d2 = w0 - 2u * w1 + w2, d3 = w0 - 3u * w1 + 3u * w2 - w3, d4 = w0 - 4u * w1 + 6u * w2 - 4u * w3 + w4;
    return (i32)(oo == 0 ? w0 : oo == 1 ? d1 : oo == 2 ? d2 : oo == 3 ? d3 : d4);
}
// calc_residuals_int (lpc.rs:279-298) at index i, in the reference's own i64 arithmetic
__device__ __forceinline__ i32 lpc_residual_at(const ChanState &cs, const LaneWin &win, int o, int i) {
    const i32 x = win.w[0];
    if (i < o) return x;
    i64 pred = 0;
#pragma unroll
    for (int t = 0; t < MAXORD; t++)
        if (t < o) pred += (i64)cs.qc[o - 5][t] * (i64)win.w[1 + t];
    pred >>= cs.lpc_shift[o - 5];
    return (i32)((u32)x - (u32)(i32)pred);
}

// ----------------------------------------------------------------------------
// Chunk loop of one analysis pass.  With two channels in the group the even warps take channel
// 0 and the odd warps channel 1, so a thread accumulates for one channel only and flushes once
// (per 64 rounds: the 32-bit partial sums hold at least 64 chunks).  The unrolled body sees the
// full chunks 1 .. nfull-1 only: they have real history and no warm-up samples.  Chunk 0 (lanes
// 0..15) and the < 16 samples behind the last full chunk (lanes 16..30) are added by the channel's
// last warp, one lane per sample, through the scalar functions above.
// ----------------------------------------------------------------------------
// Region of warp wi (of the W warps of a channel) over the full chunks 1 .. nfull-1: [a, b).  The packer uses
// the same regions; region 0 additionally owns chunk 0 and region W-1 the partial chunk behind the last full one.
__device__ __forceinline__ void region_chunks(int nfull, int W, int wi, int &a, int &b) {
    const int m = nfull > 1 ? nfull - 1 : 0;
    const int R = (m + W - 1) / W;
    a = 1 + wi * R;
    b = min(1 + (wi + 1) * R, nfull);
    if (a > b) a = b;
}

template <int NH, class Reset, class Body, class Tail, class Flush>
__device__ __forceinline__ void for_chunks(const Smem &s, int nch, Reset &&reset, Body &&body, Tail &&tailfn, Flush &&flush) {
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = nch == 2 ? (wid & 1) : 0;
    const int wi = nch == 2 ? (wid >> 1) : wid;
    const int W = nch == 2 ? NWARP / 2 : NWARP;
    const ChanState &cs = s.cs[c];
    const int nfull = cs.nfull;
    int ca, cb;
    region_chunks(nfull, W, wi, ca, cb);
    int base = ca;
    do {
        reset();
        const int end = min(cb, base + 32 * 64);
        for (int chunk = base + lane; chunk < end; chunk += 32) {
            i32 x[NH + CH];
            load_chunk<NH, true>(cs, chunk * CH, x);
            body(c, chunk, x);
        }
        if (end == cb) {
            // chunk 0 belongs to region 0 (lanes 0..15 of its warp), the samples behind the last full chunk to
            // region W - 1 (lanes 16..30 of its warp): they join the region's last flush
            int i = -1;
            if (wi == 0 && lane < 16 && nfull >= 1) i = lane;
            if (wi == W - 1 && lane >= 16 && lane - 16 < cs.tail) i = nfull * CH + lane - 16;
            if (i >= 0) { LaneWin win; lane_window(cs, i, win); tailfn(c, i, win); }
        }
        flush(c, wi);
        base = end;
    } while (base < cb);
}

// ---- pass 1: fixed-predictor statistics (sum|r|, OR|r|) for orders 0..NF-1 + autocorrelation (lpc.rs:213-221) ----
template <int P, int NF>
PASS_FN void pass1(Smem &s, int nch) {
    constexpr int NHF = NF > 1 ? 4 : 0;
    constexpr int NH = P > 8 ? 12 : (P > 0 ? 8 : NHF);
    u32 fsum[NF], forr[NF];             // |r| <= 2^20 for the fixed predictors: 2^24 per chunk
    double acc[P + 1];
    const int lane = threadIdx.x & 31;
    for_chunks<NH>(
        s, nch,
        [&]() {
#pragma unroll
            for (int o = 0; o < NF; o++) { fsum[o] = 0; forr[o] = 0; }
#pragma unroll
            for (int l = 0; l <= P; l++) acc[l] = 0.0;
        },
        [&](int c, int chunk, const i32 (&x)[NH + CH]) {
            if constexpr (P > 0) {
                // exact: |x| <= 2^16, so every partial sum is an integer far below 2^53; the P + 1 lag
                // accumulators are independent chains
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
            u32 pv[NF];
            fixed_chunk<NF, false>(x + (NH - NHF), [&](int j, const i32 (&d)[5]) {
#pragma unroll
                for (int o = 0; o < NF; o++) {
                    const u32 a = (u32)abs(d[o]);
                    if (j & 1) { fsum[o] += pv[o] + a; forr[o] |= pv[o] | a; }
                    else pv[o] = a;
                }
            });
        },
        [&](int c, int i, const LaneWin &win) {
#pragma unroll
            for (int o = 0; o < NF; o++) {
                const u32 a = (u32)abs(fixed_residual_at(win, o, i));
                fsum[o] += a; forr[o] |= a;
            }
            if constexpr (P > 0) {
                const double xi = (double)win.w[0];
#pragma unroll
                for (int l = 0; l <= P; l++) acc[l] = __fma_rn(xi, (double)win.w[l], acc[l]);
            }
        },
        [&](int c, int region) {
            ChanState &cs = s.cs[c];
#pragma unroll
            for (int o = 0; o < NF; o++) {
                const u64 t = warp_sum64((u64)fsum[o]);
                const u32 r = __reduce_or_sync(0xffffffffu, forr[o]);
                if (lane == 0) { atomicOr(&cs.fix_or[o], r); s.partA[c][1 + o][region] += t; }
            }
            if constexpr (P > 0) {
#pragma unroll
                for (int l = 0; l <= P; l++) {
                    const u64 t = warp_sum64((u64)__double2ll_rn(acc[l]));
                    if (lane == 0) s.partAC[c][l][region] += t;
                }
            }
        });
}

// ---- pass 2: LPC candidates: sum|r|, OR|r| and sum(w >> j) for the guessed window ----
// One chunk pass covers the orders LO..HI as sweeps of two orders each: (HI, LO), (HI-1, LO+1), ...
struct LpcStat { u32 sum, t0, t1, orr; };

template <int NH, int LO, int HI>
struct Sweeps {
    static constexpr int NO = HI - LO + 1;
    static __device__ __forceinline__ void run(const ChanState &cs, const i32 (&x)[NH + CH], LpcStat *st /* by order - LO0 */, int lo0) {
        if constexpr (LO <= HI) {
            // two orders per sweep while their taps fit the register file (13 coefficients), else one
            constexpr bool PAIR = FLO_PAIR_SWEEPS && LO < HI && HI + LO <= 13;
            constexpr int OA = HI, OB = PAIR ? LO : 0;
            constexpr int BS = PAIR ? FLO_PAIR_BS : FLO_SINGLE_BS;
            const bool oka = cs.lpc_ok[OA - 5] != 0, okb = OB > 0 && cs.lpc_ok[(OB > 0 ? OB : 5) - 5] != 0;
            if (oka || okb) {
                LpcStat a = st[OA - lo0], b = st[(OB > 0 ? OB : OA) - lo0];
                const int j0a = cs.lpc_j0[OA - 5], j0b = cs.lpc_j0[(OB > 0 ? OB : OA) - 5];
                u32 paa = 0, paw = 0, pav = 0, pba = 0, pbw = 0, pbv = 0;
                lpc_chunk2<NH, OA, OB, BS, false>(
                    x, cs.qd[OA - 5], cs.qd[(OB > 0 ? OB : OA) - 5],
                    [&](int j, i32 r) {
                        const u32 aa = (u32)abs(r);
                        const u32 ws = (aa + (u32)(r >> 31)) >> j0a;     // w = |r| - [r < 0]
                        const u32 wv = ws >> 1;
                        if (j & 1) { a.sum += paa + aa; a.orr |= paa | aa; a.t0 += paw + ws; a.t1 += pav + wv; }
                        else { paa = aa; paw = ws; pav = wv; }
                    },
                    [&](int j, i32 r) {
                        const u32 aa = (u32)abs(r);
                        const u32 ws = (aa + (u32)(r >> 31)) >> j0b;
                        const u32 wv = ws >> 1;
                        if (j & 1) { b.sum += pba + aa; b.orr |= pba | aa; b.t0 += pbw + ws; b.t1 += pbv + wv; }
                        else { pba = aa; pbw = ws; pbv = wv; }
                    });
                st[OA - lo0] = a;
                if constexpr (OB > 0) st[OB - lo0] = b;
            }
            Sweeps<NH, PAIR ? LO + 1 : LO, HI - 1>::run(cs, x, st, lo0);
        }
    }
};

// orders LO..HI (at most four) in one pass over the channel
template <int P, int LO, int HI>
PASS_FN void pass2_range(Smem &s, int nch) {
    constexpr int NO = HI - LO + 1;
    constexpr int NH = P > 8 ? 12 : 8;
    // 32-bit partial sums: only candidates with OR|r| < 2^21 are ever used (after_pass2), 2^25 per chunk
    LpcStat st[NO];
    const int lane = threadIdx.x & 31;
    for_chunks<NH>(
        s, nch,
        [&]() {
#pragma unroll
            for (int i = 0; i < NO; i++) { st[i].sum = 0; st[i].t0 = 0; st[i].t1 = 0; st[i].orr = 0; }
        },
        [&](int c, int chunk, const i32 (&x)[NH + CH]) {
            Sweeps<NH, LO, HI>::run(s.cs[c], x, st, LO);
        },
        [&](int c, int i, const LaneWin &win) {
            const ChanState &cs = s.cs[c];
#pragma unroll
            for (int O = LO; O <= HI; O++) {
                if (cs.lpc_ok[O - 5]) {
                    const i32 r = lpc_residual_at(cs, win, O, i);
                    const u32 a = (u32)abs(r);
                    const u32 ws = (a + (u32)(r >> 31)) >> cs.lpc_j0[O - 5];
                    st[O - LO].sum += a; st[O - LO].orr |= a; st[O - LO].t0 += ws; st[O - LO].t1 += ws >> 1;
                }
            }
        },
        [&](int c, int region) {
            ChanState &cs = s.cs[c];
#pragma unroll
            for (int i = 0; i < NO; i++) {
                const u64 a = warp_sum64((u64)st[i].sum), b = warp_sum64((u64)st[i].t0), d = warp_sum64((u64)st[i].t1);
                const u32 r = __reduce_or_sync(0xffffffffu, st[i].orr);
                if (lane == 0) {
                    atomicOr(&cs.l_or[LO - 5 + i], r);
                    s.partA[c][6 + LO - 5 + i][region] += a;
                    s.partT[c][LO - 5 + i][0][region] += b; s.partT[c][LO - 5 + i][1][region] += d;
                }
            }
        });
}

template <int P>
__device__ void pass2(Smem &s, int nch) {
    if constexpr (P <= 8) {
        pass2_range<P, 5, P>(s, nch);
    } else if constexpr (P == 10) {
        pass2_range<P, 5, 8>(s, nch);          // sweeps (8,5) (7,6)
        pass2_range<P, 9, 10>(s, nch);         // sweep (10,9)
    } else {
        pass2_range<P, 5, 8>(s, nch);
        pass2_range<P, 9, 12>(s, nch);
    }
}

// residuals of one chunk for candidate MODE (0..4 fixed, 5..12 LPC, 13 raw samples); x: NHX history samples + CH
template <int MODE, bool FIRST, int NHX, class F>
__device__ __forceinline__ void cand_chunk(const i32 (&x)[NHX + CH], const double *qd, F &&fn) {
    if constexpr (MODE == 13 || MODE == 0) {
#pragma unroll
        for (int j = 0; j < CH; j++) fn(j, x[NHX + j]);
    } else if constexpr (MODE <= 4) {
        static_assert(NHX >= 4, "history");
        fixed_chunk<MODE + 1, FIRST>(x + (NHX - 4), [&](int j, const i32 (&d)[5]) { fn(j, d[MODE]); });
    } else {
        static_assert(NHX >= MODE, "history");
        lpc_chunk2<NHX, MODE, 0, FLO_PACK_BS, FIRST>(x, qd, qd, fn, [](int, i32) {});
    }
}

// ---- pass 3: exact max|r| and S = sum(w >> j) for one still-open candidate per channel ----
template <int P, int NF>
PASS_FN void pass3(Smem &s, int nch) {
    constexpr int NH = P > 8 ? 12 : (P > 0 ? 8 : (NF > 1 ? 4 : 0));
    u64 S;
    u32 mx;
    u32 S5[NF];                        // 32-bit partial sums: (w >> j) < 2^21 per sample, flushed every 64 chunks
    const int lane = threadIdx.x & 31;
    for_chunks<NH>(
        s, nch,
        [&]() {
            S = 0; mx = 0;
#pragma unroll
            for (int o = 0; o < NF; o++) S5[o] = 0;
        },
        [&](int c, int chunk, const i32 (&x)[NH + CH]) {
            const ChanState &cs = s.cs[c];
            const int cand = cs.ex_cand;
            if (cand < 0) return;
            if (cs.ex_fixed) {
                // all open fixed candidates share one difference chain (lpc.rs:301-359); S_o for unevaluated
                // orders is computed too and simply not used
                int jj[NF];
#pragma unroll
                for (int o = 0; o < NF; o++) jj[o] = max(cs.cand_k[1 + o] - 1, 0);
                constexpr int H = NF > 1 ? 4 : 0;
                fixed_chunk<NF, false>(x + (NH - H), [&](int j, const i32 (&d)[5]) {
#pragma unroll
                    for (int o = 0; o < NF; o++) S5[o] += ((u32)abs(d[o]) + (u32)(d[o] >> 31)) >> jj[o];
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
            if constexpr (NF > 1) {
                switch (mode) {
                    case 0: cand_chunk<0, false, NH>(x, qd, fn); break;
                    case 1: cand_chunk<1, false, NH>(x, qd, fn); break;
                    case 2: cand_chunk<2, false, NH>(x, qd, fn); break;
                    case 3: if constexpr (NF > 3) cand_chunk<3, false, NH>(x, qd, fn); break;
                    case 4: if constexpr (NF > 4) cand_chunk<4, false, NH>(x, qd, fn); break;
                    default:
                        if constexpr (P > 0) {
                            ForOrders<5, P>::run([&](auto oc) {
                                constexpr int O = decltype(oc)::value;
                                if (mode == O) cand_chunk<O, false, NH>(x, qd, fn);
                            });
                        }
                        break;
                }
            } else {
                cand_chunk<0, false, NH>(x, qd, fn);
            }
            S += acc;
        },
        [&](int c, int i, const LaneWin &win) {
            const ChanState &cs = s.cs[c];
            const int cand = cs.ex_cand;
            if (cand < 0) return;
            if (cs.ex_fixed) {
#pragma unroll
                for (int o = 0; o < NF; o++) {
                    const i32 r = fixed_residual_at(win, o, i);
                    S5[o] += ((u32)abs(r) + (u32)(r >> 31)) >> max(cs.cand_k[1 + o] - 1, 0);
                }
                return;
            }
            const int k = cs.cand_k[cand];
            const int jj = k >= 1 ? k - 1 : 0;
            const int mode = cand - 1;
            const i32 r = mode <= 4 ? fixed_residual_at(win, mode, i) : lpc_residual_at(cs, win, mode, i);
            const u32 a = (u32)abs(r);
            mx = max(mx, a);
            S += (a + (u32)(r >> 31)) >> jj;
        },
        [&](int c, int region) {
            ChanState &cs = s.cs[c];
            const u64 t = warp_sum64(S);
            const u32 m = __reduce_max_sync(0xffffffffu, mx);
            if (lane == 0 && cs.ex_cand >= 0) {
                atomicMax(&cs.ex_max, m);
                if (!cs.ex_fixed) s.partS[c][cs.ex_cand][region] += t;
            }
            if (cs.ex_fixed) {
#pragma unroll
                for (int o = 0; o < NF; o++) {
                    const u64 t5 = warp_sum64((u64)S5[o]);
                    if (lane == 0 && (cs.ex_fixed >> o & 1)) s.partS[c][1 + o][region] += t5;
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
__device__ void after_pass1_warp(Smem &s, int c, int fmax, bool lpc_on) {
    ChanState &cs = s.cs[c];
    const int lane = threadIdx.x & 31;
    const u32 n = (u32)cs.n;
    // channel totals = sums over the regions
    if (lane < 5) { u64 t = 0; for (int w = 0; w < NWARP; w++) t += s.partA[c][1 + lane][w]; cs.fix_sum[lane] = t; }
    if (lane <= P) { u64 t = 0; for (int w = 0; w < NWARP; w++) t += s.partAC[c][lane][w]; cs.ac[lane] = (i64)t; }
    __syncwarp();
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
        cs.cand_src[lane] = 0;
    }
    if (lane < NLPC) cs.lpc_ok[lane] = 0;
    __syncwarp();
    if (lpc_on && cs.n > 5) {
        if constexpr (P > 0) levinson_all_orders<P>(cs);
        if (lane < NLPC && cs.n <= 5 + lane) cs.lpc_ok[lane] = 0;                 // encoder.rs:255-257
    }
    __syncwarp();
}

// after pass 2 (lanes 6..13): resolve the LPC candidates (encoder.rs:262-286).  When the Rice parameter of a
// candidate fell outside the guessed shift window, pass 2 runs a second time with the window moved onto it
// (`redo` true: the windows and the zeroed sums of the channel are set up for that run; the k of a candidate does
// not depend on the window, so the second run always hits).  Returns whether the channel asks for the second run.
__device__ bool after_pass2_warp(Smem &s, int c, int P, bool redo, u32 *counters) {
    ChanState &cs = s.cs[c];
    const int lane = threadIdx.x & 31;
    const u32 n = (u32)cs.n;
    bool hit = false, miss = false;
    const int o = lane - 1;
    const bool mine = lane >= 6 && o <= P && cs.lpc_ok[o - 5];
    int k = 0;
    if (mine) {
        const int i = o - 5;
        u64 ta = 0, t0 = 0, t1 = 0;                                               // channel totals = sums over the regions
        for (int w = 0; w < NWARP; w++) { ta += s.partA[c][lane][w]; t0 += s.partT[c][i][0][w]; t1 += s.partT[c][i][1][w]; }
        cs.l_sum[i] = ta; cs.l_t0[i] = t0; cs.l_t1[i] = t1;
        const u32 orr = cs.l_or[i];
        const int bl = bitlen32(orr);
        if (bl < 21) {                                                            // else max|r| >= 2^20 > 1_000_000: rejected
            k = rice_k_or(orr, cs.l_sum[i], n);
            cs.cand_k[lane] = k;
            cs.cand_sumabs[lane] = cs.l_sum[i];
            const int jj = k >= 1 ? k - 1 : 0;
            if (bl <= 19 && (jj == cs.lpc_j0[i] || jj == cs.lpc_j0[i] + 1)) {
                const u64 S = jj == cs.lpc_j0[i] ? cs.l_t0[i] : cs.l_t1[i];
                cs.cand_src[lane] = jj == cs.lpc_j0[i] ? 1 : 2;
                cs.cand_state[lane] = CS_EXACT;
                cs.cand_size[lane] = rice_bytes(S, cs.l_sum[i], n, k);
                hit = true;
            } else {
                cs.cand_state[lane] = CS_BOUNDED;                                 // window miss or 2^19 <= max|r| < 2^20
                miss = bl <= 19;
            }
        }
    }
    const u32 hm = __ballot_sync(0xffffffffu, hit), mm = __ballot_sync(0xffffffffu, miss);
    if (lane == 0) {
        if (hm) atomicAdd(counters + 2, (u32)__popc(hm));
        if (mm) atomicAdd(counters + 3, (u32)__popc(mm));
    }
    const bool again = redo && mm != 0;
    if (again && mine) {
        const int i = o - 5;
        if (miss) cs.lpc_j0[i] = k >= 1 ? k - 1 : 0;
        cs.l_sum[i] = 0; cs.l_or[i] = 0; cs.l_t0[i] = 0; cs.l_t1[i] = 0;
        cs.cand_state[lane] = CS_ABSENT;
        for (int w = 0; w < NWARP; w++) { s.partA[c][lane][w] = 0; s.partT[c][i][0][w] = 0; s.partT[c][i][1][w] = 0; }
    }
    __syncwarp();
    return again;
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

__device__ void after_pass3(Smem &s, int ch) {
    ChanState &cs = s.cs[ch];
    const int c = cs.ex_cand;
    if (c < 0) return;
    const u32 n = (u32)cs.n;
    if (cs.ex_fixed) {
        for (int o = 0; o < 5; o++)
            if (cs.ex_fixed >> o & 1) {
                u64 t = 0;
                for (int w = 0; w < NWARP; w++) t += s.partS[ch][1 + o][w];
                cs.ex_s5[o] = t;
                cs.cand_state[1 + o] = CS_EXACT;
                cs.cand_size[1 + o] = rice_bytes(cs.ex_s5[o], cs.cand_sumabs[1 + o], n, cs.cand_k[1 + o]);
            }
        return;
    }
    { u64 t = 0; for (int w = 0; w < NWARP; w++) t += s.partS[ch][c][w]; cs.ex_s = t; }
    if (c >= 6 && cs.ex_max > 1000000u) { cs.cand_state[c] = CS_ABSENT; return; }   // encoder.rs:269-272
    cs.cand_state[c] = CS_EXACT;
    cs.cand_size[c] = rice_bytes(cs.ex_s, cs.cand_sumabs[c], n, cs.cand_k[c]);
}

// ----------------------------------------------------------------------------
// bit packer
// ----------------------------------------------------------------------------
// The packer works in rounds of NT chunks.  A round's bit stream is staged in a ring of 32-bit words in
// shared memory: ring slot 0 stands for word `wbase` of the channel's stream (a multiple of four, and word w
// of the stream is the 4 bytes at abase + 4 w of the output with abase 16-byte aligned), so complete groups
// of four words leave as one 16-byte store and the at most three complete words plus the partial word
// behind them are moved to the front of the ring for the next round.

// Codes of one chunk for the winner: zigzag(r) (rice.rs:96), or the two little-endian bytes of
// the sample as one 16-bit MSB-first code for a raw channel ((s as i16).to_le_bytes(), encoder.rs:222-224).
template <int MODE, bool FIRST, int NHX>
__device__ __forceinline__ void chunk_codes(const ChanState &cs, int i0, const double *qd, u32 (&u)[CH]) {
    i32 x[NHX + CH];
    load_chunk<NHX, !FIRST>(cs, i0, x);
    cand_chunk<MODE, FIRST, NHX>(x, qd, [&](int j, i32 r) {
        if constexpr (MODE == 13) {
            const u32 v = (u32)r & 0xffffu;
            u[j] = ((v & 0xff) << 8) | (v >> 8);
        } else {
            u[j] = ((u32)r << 1) ^ (u32)(r >> 31);
        }
    });
}
// history the packer needs for a winner of kind MODE
template <int MODE> struct ModeHist { static constexpr int NH = (MODE == 13 || MODE == 0) ? 0 : (MODE <= 4 ? 4 : (MODE <= 8 ? 8 : 12)); };
template <int MODE>
__device__ __forceinline__ void chunk_codes_at(const ChanState &cs, int i0, const double *qd, u32 (&u)[CH]) {
    if (i0 == 0) chunk_codes<MODE, true, ModeHist<MODE>::NH>(cs, i0, qd, u);
    else chunk_codes<MODE, false, ModeHist<MODE>::NH>(cs, i0, qd, u);
}

// Appends the codes of one chunk to the staging ring, MSB first (BitWriter, rice.rs:162-208): the common
// case.  Every code is at most 31 bits (raw: 16; Rice with quotient <= 15), fewer than 32 bits are pending
// when a code is appended, so one predicated word leaves per code at most.  The first word of the chunk's
// stream is shared with the threads in front: it is parked in this thread's `headw` slot (the store address
// is swapped to the ring after the first word) and OR-ed into the ring at the end with the last, partial word.
template <bool MASKED>
__device__ __forceinline__ void emit_fast(u32 ring_addr, u32 wbase, u32 head_addr, const u32 (&u)[CH], int nv, int k, bool raw,
                                          u64 start) {
    const u32 w0 = (u32)(start >> 5);
    int nb = (int)(start & 31);                    // leading zero bits stand for the part of w0 that is not ours
    u64 acc = 0;
    u32 addr = head_addr;
    u32 next = ring_addr + ((w0 + 1u - wbase) << 2);
    auto put = [&](u32 v, int len) {               // 1 <= len <= 31, v < 2^len
        acc = (acc << len) | v;
        nb += len;
        if (nb >= 32) {
            st_shared_u32(addr, __funnelshift_r((u32)acc, (u32)(acc >> 32), nb));   // bits [nb-32, nb) of acc
            addr = next;
            next += 4;
        }
        nb &= 31;
    };
    if (raw) {
#pragma unroll
        for (int j = 0; j < CH; j++)
            if (!MASKED || j < nv) put(u[j], 16);
    } else {
        const u32 kmask = (1u << k) - 1u;
#pragma unroll
        for (int j = 0; j < CH; j++) {             // encode_sample, rice.rs:94-114
            if (!MASKED || j < nv) {
                const u32 q = u[j] >> k;
                put((((1u << q) - 1u) << (k + 1)) | (u[j] & kmask), (int)q + k + 1);
            }
        }
    }
    const u32 tailw = nb > 0 ? (u32)acc << (32 - nb) : 0u;
    const u32 slot0 = ring_addr + ((w0 - wbase) << 2);
    if (addr != head_addr) {                       // at least one word was completed: the first one is parked
        const u32 head = ld_shared_u32(head_addr);
        if (head) atom_or_shared(slot0, head);
        if (tailw) atom_or_shared(addr, tailw);
    } else if (tailw) {
        atom_or_shared(slot0, tailw);
    }
}

// General form: quotients up to 255 (rice.rs:101-106 caps at 255; codes longer than 32 bits are split), and
// only the words inside the window [wlo, whi) are written (rounds larger than the ring are emitted window
// by window; re-emitting a word is idempotent).
__device__ __noinline__ void emit_slow(u32 *ring, u32 wbase, const u32 *u, int nv, int k, bool raw, u64 start, u32 wlo, u32 whi) {
    const u32 w0 = (u32)(start >> 5);
    u32 w = w0;
    int nb = (int)(start & 31);
    u64 acc = 0;
    u32 head = 0;
    auto put = [&](u32 v, int len) {               // 1 <= len <= 32, v < 2^len
        acc = (acc << len) | v;
        nb += len;
        const bool full = nb >= 32;
        const u32 word = (u32)(acc >> ((nb - 32) & 63));
        const bool mine = full && w != w0 && w >= wlo && w < whi;
        if (mine) ring[w - wbase] = word;
        head = (full && w == w0) ? word : head;
        w += full ? 1u : 0u;
        nb -= full ? 32 : 0;
    };
    const u32 kmask = (1u << k) - 1u;
#pragma unroll 1
    for (int j = 0; j < nv; j++) {
        u32 uj = u[0];
#pragma unroll
        for (int t = 1; t < CH; t++) uj = (t == j) ? u[t] : uj;
        if (raw) { put(uj, 16); continue; }
        u32 q = uj >> k;
        while (q > 15) { const u32 t = q < 24 ? q : 24; put((1u << t) - 1u, (int)t); q -= t; }
        put((((1u << q) - 1u) << (k + 1)) | (uj & kmask), (int)q + k + 1);
    }
    const u32 tailw = nb > 0 ? (u32)(acc << (32 - nb)) : 0u;
    const bool in0 = w0 >= wlo && w0 < whi;
    if (w == w0) {
        if (in0 && tailw) atomicOr(&ring[w0 - wbase], tailw);
    } else {
        if (in0 && head) atomicOr(&ring[w0 - wbase], head);
        if (tailw && w >= wlo && w < whi) atomicOr(&ring[w - wbase], tailw);
    }
}

__device__ __forceinline__ u32 bswap32(u32 v) { return __byte_perm(v, 0, 0x0123); }

// What a warp needs to know about the bytes it may write: [lo, hi) are the bytes of the words it owns alone;
// the word index `ehw` / `etw` (when `eh` / `et`) is shared with the region in front / behind, and its bits are
// collected in the CTA's edge accumulators instead (shared addresses ehp / etp).
struct RegionOut {
    uint8_t *obase;
    u64 abase, lo, hi;
    u32 ehw, etw, ehp, etp;
    bool eh, et;
};
__device__ __forceinline__ void store_word_masked(const RegionOut &o, u32 w, u32 v) {
    const u64 a = o.abase + 4ull * w;
#pragma unroll
    for (int b = 0; b < 4; b++)
        if (a + b >= o.lo && a + b < o.hi) o.obase[a + b] = (uint8_t)(v >> (24 - 8 * b));
    if (o.eh && w == o.ehw && v) atom_or_shared(o.ehp, v);
    if (o.et && w == o.etw && v) atom_or_shared(o.etp, v);
}
// Write `ng` complete groups of four ring words (slots 0 .. 4 ng) of this warp's ring to the output and clear
// them; then lane 0 moves the group behind them (complete words that do not fill a group yet + the partial
// word) to slot 0.  Group g is the 16 bytes at abase + 4 wbase + 16 g.
__device__ __forceinline__ void flush_groups(u32 *ring, const RegionOut &o, u32 wbase, u32 ng) {
    const u32 lane = threadIdx.x & 31;
    uint4 *r4 = reinterpret_cast<uint4 *>(ring);
    for (u32 g = lane; g < ng; g += 32) {
        const uint4 v = r4[g];
        r4[g] = make_uint4(0, 0, 0, 0);
        const u64 a = o.abase + 4ull * wbase + 16ull * g;
        if (a >= o.lo && a + 16 <= o.hi) {
            *reinterpret_cast<uint4 *>(o.obase + a) = make_uint4(bswap32(v.x), bswap32(v.y), bswap32(v.z), bswap32(v.w));
        } else {
            store_word_masked(o, wbase + 4 * g, v.x); store_word_masked(o, wbase + 4 * g + 1, v.y);
            store_word_masked(o, wbase + 4 * g + 2, v.z); store_word_masked(o, wbase + 4 * g + 3, v.w);
        }
    }
    if (lane == 0 && ng > 0) {
        const uint4 c = r4[ng];
        r4[ng] = make_uint4(0, 0, 0, 0);
        r4[0] = c;
    }
}

// samples of region wi: its full chunks, chunk 0 for region 0, the partial chunk for region W - 1
__device__ __forceinline__ u32 region_samples(int n, int W, int wi) {
    const int nfull = n / CH, tail = n % CH;
    int ca, cb;
    region_chunks(nfull, W, wi, ca, cb);
    return (u32)(cb - ca) * CH + ((wi == 0 && nfull >= 1) ? (u32)CH : 0u) + (wi == W - 1 ? (u32)tail : 0u);
}

// Pack the residual payload of region wi (of W) of one channel; called by one warp.  The payload starts at
// byte offset `pos` (relative to obase, which is 16-byte aligned) -- encode_i32 / BitWriter (rice.rs:84-92,
// 162-208) or encode_raw (encoder.rs:220-226).  The region's bit offset inside the payload is the sum of the
// bit counts of the regions in front (cr.regbits, from the analysis passes); no block-wide step is needed.
template <int P>
PASS_FN void pack_region(Smem &s, u32 *ring, const ChanState &cs, int cq, const ChanResult &cr, uint8_t *obase, u64 pos, int wi, int W,
                            u32 *err) {
    const int lane = threadIdx.x & 31;
    const int n = cs.n, nfull = n / CH, tail = n % CH;
    const int mode = cr.kind == 0 ? 13 : cr.order;
    const bool raw = cr.kind == 0;
    const int k = cr.k;
    u64 bitstart = (pos & 15ull) * 8ull;
    for (int v = 0; v < wi; v++) bitstart += cr.regbits[v];
    const u64 bits = cr.regbits[wi];
    if (bits == 0) return;
    const u64 bitend = bitstart + bits;
    RegionOut o;
    o.obase = obase;
    o.abase = pos & ~15ull;
    o.eh = wi > 0 && (bitstart & 31) != 0;
    o.et = wi < W - 1 && (bitend & 31) != 0;
    o.ehw = (u32)(bitstart >> 5); o.etw = (u32)(bitend >> 5);
    o.ehp = smem_u32(&s.edge[cq][wi]); o.etp = smem_u32(&s.edge[cq][wi + 1]);
    o.lo = wi == 0 ? pos : o.abase + 4ull * ((bitstart + 31) >> 5);
    o.hi = wi == W - 1 ? pos + cr.nbytes : o.abase + 4ull * (bitend >> 5);
    int ca, cb;
    region_chunks(nfull, W, wi, ca, cb);
    const int cstart = (wi == 0 && nfull >= 1) ? 0 : ca;
    const int cend = cb + ((wi == W - 1 && tail > 0) ? 1 : 0);
    const double *qd = s.wqd[cq];
    const u32 ring_addr = smem_u32(ring), head_addr = smem_u32(&s.headw[threadIdx.x]);
    u64 bitpos = bitstart;
    u32 wbase = (u32)(bitstart >> 5) & ~3u;            // ring slot 0 <-> stream word wbase (multiple of 4)
    for (int base = cstart; base < cend; base += 32) {
        const int chunk = base + lane;
        const int i0 = chunk * CH;
        const int nv = chunk < cend ? (chunk == nfull ? tail : CH) : 0;
        u32 u[CH];
        u32 tb = 0;
        bool fast = true;
        if (nv > 0) {
            switch (mode) {
                case 0: chunk_codes_at<0>(cs, i0, qd, u); break;
                case 1: chunk_codes_at<1>(cs, i0, qd, u); break;
                case 2: chunk_codes_at<2>(cs, i0, qd, u); break;
                case 3: chunk_codes_at<3>(cs, i0, qd, u); break;
                case 4: chunk_codes_at<4>(cs, i0, qd, u); break;
                case 13: chunk_codes_at<13>(cs, i0, qd, u); break;
                default:
                    if constexpr (P > 0) {
                        ForOrders<5, P>::run([&](auto oc) {
                            constexpr int O = decltype(oc)::value;
                            if (mode == O) chunk_codes_at<O>(cs, i0, qd, u);
                        });
                    }
                    break;
            }
            if (raw) {
                tb = 16u * (u32)nv;
            } else if (nv == CH) {
                u32 qs = 0, orr = 0;
#pragma unroll
                for (int j = 0; j < CH; j += 2) { qs += (u[j] >> k) + (u[j + 1] >> k); orr |= u[j] | u[j + 1]; }
                tb = qs + (u32)CH * (1u + (u32)k);
                fast = (orr >> k) <= 15u;              // every quotient <= 15: codes of at most 31 bits
            } else {
                u32 orr = 0;
#pragma unroll
                for (int j = 0; j < CH; j++)
                    if (j < nv) { tb += (u[j] >> k) + 1u + (u32)k; orr |= u[j]; }
                fast = (orr >> k) <= 15u;
            }
        }
        // warp exclusive scan of the chunk bit counts
        u32 inc = tb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        const u32 total = __shfl_sync(0xffffffffu, inc, 31);
        const u64 start = bitpos + (inc - tb);
        const u64 end_r = bitpos + total;
        const u32 wlast = (u32)((end_r + 31) >> 5);
        if (wlast - wbase <= (u32)WRING - 4u) {        // (the group behind the last complete one is moved, too)
            if (nv > 0) {
                if (fast) {
                    if (nv == CH) emit_fast<false>(ring_addr, wbase, head_addr, u, CH, k, raw, start);
                    else emit_fast<true>(ring_addr, wbase, head_addr, u, nv, k, raw, start);
                } else {
                    u32 uc[CH];                    // (a copy: keeps u itself in registers on the common path)
#pragma unroll
                    for (int j = 0; j < CH; j++) uc[j] = u[j];
                    emit_slow(ring, wbase, uc, nv, k, raw, start, wbase, wbase + WRING - 4u);
                }
            }
            __syncwarp();
            const u32 ng = ((u32)(end_r >> 5) - wbase) >> 2;
            flush_groups(ring, o, wbase, ng);
            wbase += 4u * ng;
            __syncwarp();
        } else {
            u32 uc[CH];
#pragma unroll
            for (int j = 0; j < CH; j++) uc[j] = u[j];
            for (;;) {
                const u32 whi = wbase + WRING - 4u;
                if (nv > 0) emit_slow(ring, wbase, uc, nv, k, raw, start, wbase, whi);
                __syncwarp();
                const u32 wend = min(whi, (u32)(end_r >> 5));
                const u32 ng = (wend - wbase) >> 2;
                flush_groups(ring, o, wbase, ng);
                wbase += 4u * ng;
                __syncwarp();
                if (whi >= wlast) break;
            }
        }
        bitpos = end_r;
    }
    {
        // what is left: at most three complete words and the partial one
        const u32 wfin = (u32)((bitpos + 31) >> 5);
        if (wbase + (u32)lane < wfin) {
            const u32 v = ring[lane];
            ring[lane] = 0;
            store_word_masked(o, wbase + (u32)lane, v);
        }
    }
    if (lane == 0 && bitpos != bitend) atomicExch(err, 0xBAD00003u);
    __syncwarp();
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

// Two halves, both called by warp 0: publish makes this frame's size visible to its successors, wait returns
// the exclusive prefix in every lane and publishes the inclusive one.  Whatever the CTA does between the two
// (packing a small frame into its scratch) is time its predecessors have to publish in.
__device__ __forceinline__ void lookback_publish(u64 *status, u32 g, u64 mine) {
    if ((threadIdx.x & 31) == 0) st_status(status + g, (g == 0 ? ST_PRE : ST_AGG) | mine);
}
__device__ u64 lookback_wait(u64 *status, u32 g, u64 mine) {
    const int lane = threadIdx.x & 31;
    if (g == 0) return 0;
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
__device__ __forceinline__ u64 lookback_exclusive(u64 *status, u32 g, u64 mine) {
    lookback_publish(status, g, mine);
    return lookback_wait(status, g, mine);
}

// A frame packed into the CTA's scratch (16-byte aligned, frame-relative) goes to its place in the output, which
// may start at any byte: whole 16-byte units of the destination are assembled from five aligned source words,
// the <= 15 bytes in front of and behind them are stored byte-wise (the neighbours belong to other CTAs).
__device__ __forceinline__ void copy_frame_out(const uint8_t *src, uint8_t *dst, u32 n) {
    const u32 tid = threadIdx.x;
    const u32 head = min(n, (u32)((16u - (u32)((uintptr_t)dst & 15u)) & 15u));
    if (tid < head) dst[tid] = __ldcg(src + tid);
    const u32 units = (n - head) >> 4;
    const u32 sb = 8u * (head & 3u);
    const u32 *s32 = reinterpret_cast<const u32 *>(src) + (head >> 2);
    uint4 *d16 = reinterpret_cast<uint4 *>(dst + head);
#pragma unroll 4
    for (u32 u = tid; u < units; u += NT) {
        const u32 *q = s32 + 4 * u;
        const u32 w0 = __ldcg(q), w1 = __ldcg(q + 1), w2 = __ldcg(q + 2), w3 = __ldcg(q + 3), w4 = sb ? __ldcg(q + 4) : 0u;
        d16[u] = make_uint4(__funnelshift_r(w0, w1, sb), __funnelshift_r(w1, w2, sb), __funnelshift_r(w2, w3, sb), __funnelshift_r(w3, w4, sb));
    }
    const u32 done = head + 16u * units;
    if (tid < n - done) dst[done + tid] = __ldcg(src + done + tid);
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

// One group of SPT sample frames (C = 1 or 2 channels) from a staged piece of the interleaved input in
// shared memory: quantise (a3), silence test (a4), deinterleave (a5), energies (a6), plane stores.
template <typename T, int C>
__device__ __forceinline__ void ingest_group(const unsigned char *src, IngestAcc &A, int16_t *planes, u32 stride, u32 grp) {
    constexpr int NS = SPT * C;                        // samples of the group
    constexpr int NB = NS * (int)sizeof(T);            // bytes: 8 .. 64
    static_assert(NB % 8 == 0, "group bytes");
    u32 raw[NB / 4];
    if constexpr (NB % 16 == 0) {
#pragma unroll
        for (int i = 0; i < NB / 16; i++) {
            const uint4 v = reinterpret_cast<const uint4 *>(src)[i];
            raw[4 * i] = v.x; raw[4 * i + 1] = v.y; raw[4 * i + 2] = v.z; raw[4 * i + 3] = v.w;
        }
    } else {
        const uint2 v = *reinterpret_cast<const uint2 *>(src);
        raw[0] = v.x; raw[1] = v.y;
    }
    float f[NS];
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int i = 0; i < NS; i++) f[i] = __uint_as_float(raw[i]);
    } else {
#pragma unroll
        for (int i = 0; i < NS / 2; i++) {
            f[2 * i] = pcm_to_f32((int)(int16_t)(raw[i] & 0xffffu));
            f[2 * i + 1] = pcm_to_f32((int)raw[i] >> 16);
        }
    }
    u32 w0[SPT / 2], w1[SPT / 2];
    if constexpr (C == 2) {
#pragma unroll
        for (int i = 0; i < SPT / 2; i++) {
            i32 la, ra, lb, rb;
            A.pair(f[4 * i], f[4 * i + 1], la, ra);
            A.pair(f[4 * i + 2], f[4 * i + 3], lb, rb);
            w0[i] = pack2(la, lb); w1[i] = pack2(ra, rb);
        }
    } else {
#pragma unroll
        for (int i = 0; i < SPT / 2; i++) {
            A.loud |= is_loud(f[2 * i]) | is_loud(f[2 * i + 1]);
            w0[i] = pack2(f32_to_i32(f[2 * i]), f32_to_i32(f[2 * i + 1]));
        }
    }
    int16_t *d0 = planes + (size_t)grp * SPT;
    if constexpr (SPT == 8) {
        *reinterpret_cast<uint4 *>(d0) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
        if constexpr (C == 2) *reinterpret_cast<uint4 *>(d0 + stride) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
    } else {
        *reinterpret_cast<uint2 *>(d0) = make_uint2(w0[0], w0[1]);
        if constexpr (C == 2) *reinterpret_cast<uint2 *>(d0 + stride) = make_uint2(w1[0], w1[1]);
    }
}

// The group-aligned part of a frame (C = 1 or 2), staged through `work` by bulk async copies: step k is the
// NT groups [k NT, (k+1) NT) -- one contiguous piece of the interleaved input -- and lands in stage k % nst.
// Thread 0 is the producer; a stage is refilled after the CTA barrier that ends its step.  Returns the number
// of sample frames done.  `phase` holds the parity of every stage's mbarrier and lives across frames.
template <typename T, int C>
__device__ __forceinline__ u32 ingest_staged(Smem &s, unsigned char *work, u32 nst, const T *in, u32 nf, IngestAcc &A,
                                             int16_t *planes, u32 stride, u32 &phase) {
    constexpr u32 BT = SPT * C * (u32)sizeof(T);       // bytes per thread and step
    constexpr u32 STAGE = NT * BT;
    const int tid = threadIdx.x;
    u32 ngrp = nf / SPT;
    if (BT % 16 != 0) ngrp &= ~1u;                     // bulk copies move multiples of 16 bytes
    const u32 nsteps = (ngrp + NT - 1) / NT;
    const u32 work_addr = smem_u32(work), bar0 = smem_u32(&s.bar_full[0]);
    auto issue = [&](u32 k) {
        const u32 st = k % nst;
        const u32 groups = min((u32)NT, ngrp - k * NT);
        const u32 bytes = groups * BT;
        mbar_expect_tx(bar0 + 8 * st, bytes);
        bulk_load(work_addr + st * STAGE, reinterpret_cast<const unsigned char *>(in) + (size_t)k * STAGE, bytes, bar0 + 8 * st);
    };
    if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the work area was last touched through the generic proxy
        for (u32 k = 0; k < min(nst, nsteps); k++) issue(k);
    }
    for (u32 k = 0; k < nsteps; k++) {
        const u32 st = k % nst;
        mbar_wait(bar0 + 8 * st, (phase >> st) & 1u);
        phase ^= 1u << st;
        const u32 grp = k * NT + tid;
        if (grp < ngrp) ingest_group<T, C>(work + st * STAGE + tid * BT, A, planes, stride, grp);
        __syncthreads();
        if (tid == 0 && k + nst < nsteps) issue(k + nst);
    }
    return ngrp * SPT;
}

template <typename T>
__device__ void ingest_frame(Smem &s, unsigned char *work, u32 work_bytes, const T *in, u32 len, u32 C, int16_t *planes, u32 stride,
                             u32 &phase) {
    const int tid = threadIdx.x;
    IngestAcc A;
    const bool aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
    if (C == 2) {
        const u32 nf = len >> 1;
        u32 done = 0;
        const u32 nst = min((u32)MAX_STAGES, work_bytes / (NT * SPT * 2u * (u32)sizeof(T)));
        if (aligned && nst >= 2) done = ingest_staged<T, 2>(s, work, nst, in, nf, A, planes, stride, phase);
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
        const u32 nst = min((u32)MAX_STAGES, work_bytes / (NT * SPT * (u32)sizeof(T)));
        if (aligned && nst >= 2) done = ingest_staged<T, 1>(s, work, nst, in, len, A, planes, stride, phase);
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

// Ask for the interleaved input of global frame g to be brought into L2 (one warp; a hint only).
__device__ __forceinline__ void prefetch_frame_l2(const EncodeParams &p, u32 g) {
    const int lane = threadIdx.x & 31;
    const uint2 fd = p.frames[g];
    const TrackDev tr = p.tracks[fd.x];
    const u64 esz = p.format == FLO_FMT_PCM16 ? 2 : 4;
    const u64 spf = (u64)tr.sample_rate * tr.channels;
    const u64 b0 = (u64)fd.y * spf, b1 = min(b0 + spf, tr.n_inter);
    uintptr_t a0 = (reinterpret_cast<uintptr_t>(tr.samples) + b0 * esz + 15) & ~(uintptr_t)15;
    const uintptr_t a1 = (reinterpret_cast<uintptr_t>(tr.samples) + b1 * esz) & ~(uintptr_t)15;
    if (a1 <= a0) return;
    const u64 piece = (((a1 - a0) / 32) + 15) & ~15ull;
    const uintptr_t q0 = a0 + piece * lane;
    if (piece == 0 || q0 >= a1) return;
    const u64 nb = min((u64)piece, (u64)(a1 - q0));
    l2_prefetch_bulk(reinterpret_cast<const void *>(q0), (u32)nb);
}

// ----------------------------------------------------------------------------
// CRC32 of the frame just written (crc32.rs:23-30; the DATA chunk's CRC is folded from the frames' by
// k_crc_frames).  Raw CRC R(M) = M(x) x^32 mod p: linear, ignores leading zeros.  Thread t of the CTA reads the
// aligned words t, t + NT, ... of the frame (coalesced; the bytes were written by this CTA a moment ago and come
// from L2, eight loads in flight per thread) and runs d = d x^(32 NT) + w -- four table look-ups, like
// slice-by-4.  The words are aligned to the END of the frame, so thread t's stream weighs x^(32 (NT - t)) whatever
// the frame's length; the XOR over the threads is the state after the last whole word, and the <= 3 bytes behind
// it are added byte-wise.  multmodp follows zlib's crc32_combine helper (zlib 1.2.12+, crc32.c; (C) 1995-2022
// Mark Adler, zlib licence), re-typed for the reflected polynomial.
// ----------------------------------------------------------------------------
__device__ __forceinline__ u32 crc_multmodp(u32 a, u32 b) {
    u32 m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) { p ^= b; if ((a & (m - 1)) == 0) break; }
        m >>= 1;
        b = (b & 1) ? (b >> 1) ^ 0xEDB88320u : b >> 1;
    }
    return p;
}
// raw CRC of the bytes [a, a + fsize) of out -> p.frame_crc[g]; the bytes must be visible (barrier before)
__device__ __noinline__ void crc_frame(Smem &s, const EncodeParams &p, const uint8_t *out, u32 g, u64 a, u32 fsize) {
    const int tid = threadIdx.x;
    const u64 b = a + fsize, b4 = b & ~3ull;
    if (b4 > a) {
        const u64 a4 = a & ~3ull;
        const u64 nwords = (b4 - a4) >> 2;
        const u64 J = (nwords + NT - 1) / NT;
        const i64 s0 = (i64)b4 - (i64)(4ull * NT * J);       // may lie before a4: those words count as zero
        const u32 head_mask = 0xFFFFFFFFu << (8 * (u32)(a - a4));
        u32 d = 0;
        i64 addr = s0 + 4 * (i64)tid;
        for (u64 j = 0; j < J; j += 8) {
            u32 w[8];
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const i64 aq = addr + (i64)q * 4 * NT;
                w[q] = (j + q < J && aq >= (i64)a4) ? __ldcg(reinterpret_cast<const u32 *>(out + aq)) : 0u;
                if (aq == (i64)a4) w[q] &= head_mask;
            }
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (j + q < J) d = s.crc_x[0][d & 0xff] ^ s.crc_x[1][(d >> 8) & 0xff] ^ s.crc_x[2][(d >> 16) & 0xff] ^ s.crc_x[3][d >> 24] ^ w[q];
            addr += 8ll * 4 * NT;
        }
        u32 c = crc_multmodp(s.crc_klane[tid], d);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, o);
        if ((tid & 31) == 0 && c) atomicXor(&s.fcrc, c);
    }
    __syncthreads();
    if (tid == 0) {
        u32 c = s.fcrc;
        for (u64 i = b4 > a ? b4 : a; i < b; i++) c = (c >> 8) ^ s.crc_slice0[(c ^ __ldcg(out + i)) & 0xff];
        p.frame_crc[g] = c;
        s.fcrc = 0;
    }
}

static_assert(sizeof(ChanResult) % 4 == 0, "ChanResult is copied word by word");
// ----------------------------------------------------------------------------
// the frame-encode kernel
// ----------------------------------------------------------------------------
extern __shared__ __align__(128) unsigned char dyn_smem[];

#ifndef FLO_LB_THREADS
#define FLO_LB_THREADS NT        // experiments: a larger value lowers the register bound without changing the launch
#endif
template <int P>
__global__ void __launch_bounds__(FLO_LB_THREADS, P == 0 ? FLO_VARIANT_CTAS_FIXED : FLO_VARIANT_CTAS) k_encode_frames(const EncodeParams p) {
    Smem &s = *reinterpret_cast<Smem *>(dyn_smem);
    unsigned char *work = dyn_smem + SMEM_HDR;
    u32 *ring = reinterpret_cast<u32 *>(work);
    int16_t *smem_planes = reinterpret_cast<int16_t *>(work + p.work_bytes);
    const int tid = threadIdx.x;
    LV(g_lev_phase = p.phase_cycles;)
    if (tid < 8) s.cnt[tid] = 0;
    for (int i = tid; i < 1024; i += NT) (&s.crc_x[0][0])[i] = p.crc_tab[crc_tab_offset(NT) + i];
    s.crc_klane[tid] = p.crc_tab[crc_tab_offset(NT) + 1024 + tid];
    for (int i = tid; i < 256; i += NT) s.crc_slice0[i] = p.crc_tab[i];
    if (tid == 0) {
        s.fcrc = 0;
        for (int i = 0; i < MAX_STAGES; i++) mbar_init(smem_u32(&s.bar_full[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    u32 tma_phase = 0;
    ChanResult *cres = p.cres + (size_t)blockIdx.x * 256;
    const int level = p.level;
    // P = LPC max order analysed by this instantiation (0: levels 0-3, fixed predictors only)
    const int PL = order_of_level(level);            // encoder.rs:289-302
    const int fmax = PL < 4 ? PL : 4;
    const bool lpc_on = P > 0;                        // level >= 3 && max_order > 4, encoder.rs:204
    const bool prune = p.report == nullptr;      // the parity report wants every candidate's exact size

    if (p.stagger && blockIdx.x >= gridDim.x / 2) {
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)p.stagger) __nanosleep(200);
    }
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            // A ticket is taken only when the CTA is ready to start the frame: frames are then published
            // (look-back status) in nearly ticket order.
            s.g = p.frame_begin + atomicAdd(p.ticket, 1u);
            s.loud = 0; s.ms = 0;
            s.ms_var[0] = s.ms_var[1] = s.ms_var[2] = 0;
        }
        __syncthreads();
        const u32 g = s.g;
        if (g >= p.frame_end) break;
        PH(const long long tc0 = clock64();)
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

        PH(const long long tcA = clock64();)
        if (p.format == FLO_FMT_PCM16)
            ingest_frame<int16_t>(s, work, p.work_bytes, reinterpret_cast<const int16_t *>(tr.samples) + start, len, C, planes, stride, tma_phase);
        else
            ingest_frame<float>(s, work, p.work_bytes, reinterpret_cast<const float *>(tr.samples) + start, len, C, planes, stride, tma_phase);
        PH(const long long tcB = clock64();)
        // the work area becomes the packer's staging ring: all zero
        for (int i = tid; i < NWARP * WRING / 4; i += NT) reinterpret_cast<uint4 *>(ring)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        PH(if (tid == 0) { if (!LEVC) atomicAdd(p.phase_cycles + 12, (u64)(tcA - tc0)); atomicAdd(p.phase_cycles + 13, (u64)(tcB - tcA)); })

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
            __syncthreads();
            crc_frame(s, p, p.out, g, data_base + s.frame_excl, fsize);
            if (p.report) {
                for (u32 i = tid; i < REPORT_CH * NCAND; i += NT) {
                    flo_cand_report *r = p.report + (size_t)g * REPORT_CH * NCAND + i;
                    r->k = 0; r->pad = 0; r->size = -1;
                }
            }
            continue;
        }

        PH(const long long tc1 = clock64();)
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
            for (int i = tid; i < GROUP * NCAND * NWARP; i += NT) { (&s.partA[0][0][0])[i] = 0; (&s.partS[0][0][0])[i] = 0; }
            for (int i = tid; i < GROUP * NLPC * 2 * NWARP; i += NT) (&s.partT[0][0][0][0])[i] = 0;
            for (int i = tid; i < GROUP * (MAXORD + 1) * NWARP; i += NT) (&s.partAC[0][0][0])[i] = 0;
            if (tid < nch) {
                ChanState &cs = s.cs[tid];
                const u32 c = c0 + tid;
                u32 cl = len > c ? (len - c + C - 1) / C : 0;
                if (ms) cl = len >> 1;                                      // zip in to_mid_side truncates, encoder.rs:160-167
                cs.n = (int)cl;
                cs.nfull = (int)(cl / CH);
                cs.tail = (int)(cl % CH);
                cs.msmode = ms ? (c == 0 ? 1 : 2) : 0;
                cs.sel_lo = cs.msmode == 2 ? 0x00ff : 0x0001;
                cs.sel_hi = cs.msmode == 2 ? 0xff00 : 0x0100;
                cs.pa = ms ? planes : planes + (size_t)c * stride;
                cs.pb = planes + stride;
                cs.glob = planes != smem_planes;
                for (int o = 0; o < 5; o++) { cs.fix_sum[o] = 0; cs.fix_or[o] = 0; }
                for (int l = 0; l <= MAXORD; l++) cs.ac[l] = 0;
                for (int o = 0; o < NLPC; o++) { cs.l_sum[o] = 0; cs.l_or[o] = 0; cs.l_t0[o] = 0; cs.l_t1[o] = 0; cs.lpc_ok[o] = 0; }
                cs.ex_cand = -1; cs.ex_s = 0; cs.ex_max = 0; cs.ex_fixed = 0;
                s.redo[tid] = 0;
            }
            if (tid == 1 && nch == 1) s.redo[1] = 0;
            __syncthreads();
            PH(const long long ta0 = clock64();)
            bool any_lpc = false;
            for (int q = 0; q < nch; q++) any_lpc |= lpc_on && s.cs[q].n > 5;
            if (any_lpc) pass1<P, 5>(s, nch);
            else if (fmax >= 3) pass1<0, 5>(s, nch);
            else if (fmax == 2) pass1<0, 3>(s, nch);
            else pass1<0, 1>(s, nch);
            __syncthreads();
            PH(const long long ta1 = clock64();)
            if ((tid >> 5) < nch && s.cs[tid >> 5].n > 0) after_pass1_warp<P>(s, tid >> 5, fmax, lpc_on);
            LV(if (tid == 0) atomicAdd(p.phase_cycles + 14, (u64)(clock64() - ta1)); if (tid == 32) atomicAdd(p.phase_cycles + 12, (u64)(clock64() - ta1));)
            __syncthreads();
            PH(const long long ta2 = clock64();)
            bool run2 = false;
            for (int q = 0; q < nch; q++)
                for (int o = 0; o < NLPC; o++) run2 |= s.cs[q].n > 0 && s.cs[q].lpc_ok[o] != 0;
            PH(long long ta3 = 0;)
            for (int run = 0; run2 && run < 2; run++) {
                if constexpr (P > 0) pass2<P>(s, nch);
                __syncthreads();
                PH(if (run == 0) ta3 = clock64();)
                if ((tid >> 5) < nch && s.cs[tid >> 5].n > 0) {
                    const bool again = after_pass2_warp(s, tid >> 5, P, run == 0, s.cnt);
                    if ((tid & 31) == 0 && again) s.redo[tid >> 5] = 1;
                }
                __syncthreads();
                if (run == 1 || !(s.redo[0] | s.redo[nch - 1])) break;
                // second run of pass 2; a channel of the group that had no miss redoes its sums, too (same windows)
                if ((tid >> 5) < nch && s.cs[tid >> 5].n > 0 && !s.redo[tid >> 5]) {
                    ChanState &cs = s.cs[tid >> 5];
                    const int lane = tid & 31, i = lane - 6;
                    if (i >= 0 && i < NLPC && cs.lpc_ok[i]) {
                        cs.l_sum[i] = 0; cs.l_or[i] = 0; cs.l_t0[i] = 0; cs.l_t1[i] = 0;
                        cs.cand_state[lane] = CS_ABSENT;
                        for (int w = 0; w < NWARP; w++) { s.partA[tid >> 5][lane][w] = 0; s.partT[tid >> 5][i][0][w] = 0; s.partT[tid >> 5][i][1][w] = 0; }
                    }
                }
                __syncthreads();
            }
            PH(if (!run2) ta3 = clock64();)
            if ((tid >> 5) < nch && s.cs[tid >> 5].n > 0) next_open_candidate_warp(s.cs[tid >> 5], prune, s.cnt);
            __syncthreads();
            // exact evaluation of whatever is still open (bounded candidates that can still win)
            for (;;) {
                bool more = false;
                for (int q = 0; q < nch; q++) more |= s.cs[q].n > 0 && s.cs[q].ex_cand >= 0;
                if (!more) break;
                if (tid == 0) atomicAdd(&s.cnt[1], 1u);
                if (P > 0 || fmax >= 3) pass3<P, 5>(s, nch);
                else if (fmax == 2) pass3<0, 3>(s, nch);
                else pass3<0, 1>(s, nch);
                __syncthreads();
                if ((tid >> 5) < nch && s.cs[tid >> 5].n > 0) {
                    ChanState &cs = s.cs[tid >> 5];
                    if ((tid & 31) == 0) after_pass3(s, tid >> 5);
                    __syncwarp();
                    next_open_candidate_warp(cs, prune, s.cnt);
                }
                __syncthreads();
            }
            PH(if (tid == 0) {
                const long long ta4 = clock64();
                atomicAdd(p.phase_cycles + 8, (u64)(ta1 - ta0)); atomicAdd(p.phase_cycles + 9, (u64)(ta2 - ta1));
                atomicAdd(p.phase_cycles + 10, (u64)(ta3 - ta2)); atomicAdd(p.phase_cycles + 11, (u64)(ta4 - ta3));
            })
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
                    // payload bits of every packer region (rice.rs:97-113 summed over the region's samples)
                    const int W = nch == 2 ? NWARP / 2 : NWARP;
                    u64 allbits = 0;
                    for (int w = 0; w < 16; w++) {
                        u64 b = 0;
                        if (w < W) {
                            const u64 cnt = region_samples(cs.n, W, w);
                            if (bj == 0) b = 16ull * cnt;
                            else {
                                const int src = cs.cand_src[bj];
                                const u64 S = src == 0 ? s.partS[tid][bj][w] : s.partT[tid][bj - 6][src - 1][w];
                                b = r.k >= 1 ? S + cnt * (u64)(1 + r.k) : S + s.partA[tid][bj][w] + cnt;
                            }
                        }
                        r.regbits[w] = b;
                        allbits += b;
                    }
                    if (((allbits + 7) >> 3) != (u64)r.nbytes) atomicExch(p.err, 0xBAD00004u);
                }
                if (cs.n == 0) for (int w = 0; w < 16; w++) r.regbits[w] = 0;
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

        PH(const long long tc2 = clock64();)
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
        // A frame that fits the CTA's scratch is packed there, frame-relative, and copied to its place afterwards:
        // its offset is then only asked for after the pack, when the frames in front of it have long published
        // their sizes.  With 592 CTAs of short frames in flight the wait in front of the pack was 30 % of the
        // frame's time (every CTA waits for the slowest analysis among its predecessors, then all move on together).
        const bool defer = p.defer_bytes != 0 && fsize + 64u <= p.defer_bytes;
        if (tid < 32) {
            lookback_publish(p.status, g, fsize);
            if (!defer) {
                u64 ex = lookback_wait(p.status, g, fsize);
                if (tid == 0) { s.frame_excl = ex; p.frame_excl[g] = ex; p.frame_size[g] = fsize; }
            }
        }
#ifdef FLO_PREFETCH_NEXT_FRAME
        // Asking for the input of the frame this CTA is likely to take next (g + gridDim.x) while this one is packed
        // was measured: no gain in kernel time, and 296 frames of input parked in L2 beside the planes push each
        // other out -- DRAM reads rose from 1.0 x to 1.4 x the input bytes.  Off.
        else if (tid >= NT - 32) {
            const u32 gn = g + gridDim.x;
            if (gn < p.frame_end) prefetch_frame_l2(p, gn);
        }
#endif
        __syncthreads();

        PH(const long long tc3 = clock64();)
        // write the frame, writer.rs:236-301
        const u64 fpos = defer ? 0ull : data_base + s.frame_excl;
        uint8_t *o = defer ? p.defer_scratch + (size_t)blockIdx.x * p.defer_bytes : p.out;
        if (tid == 0) {
            o[fpos] = (uint8_t)(all_raw ? 254u : frame_type_alpc);
            put_u32le(o + fpos + 1, frame_samples);
            o[fpos + 5] = (uint8_t)(ms ? 1 : 0);
        }
        u64 pos = fpos + 6;
        const int wid = tid >> 5;
        for (u32 c0 = 0; c0 < C; c0 += GROUP) {
            // the channels of a group are packed side by side (even warps: first channel, odd warps: second one),
            // every warp its own region of the channel -- the same regions the analysis passes summed over
            const int nchp = (int)min((u32)GROUP, C - c0);
            const int cq = nchp == 2 ? (wid & 1) : 0;
            const int wi = nchp == 2 ? (wid >> 1) : wid;
            const int W = nchp == 2 ? NWARP / 2 : NWARP;
            u64 cpos[GROUP];
            u32 chdr[GROUP];
            __syncthreads();
            {   // the winners of this group: global -> shared, once (every later read is a shared-memory read)
                constexpr int WORDS = (int)(sizeof(ChanResult) / 4);
                for (int i = tid; i < nchp * WORDS; i += NT)
                    reinterpret_cast<u32 *>(&s.wres[i / WORDS])[i % WORDS] = reinterpret_cast<const u32 *>(&cres[c0 + i / WORDS])[i % WORDS];
            }
            if (tid < nchp) {
                ChanState &cs = s.cs[tid];
                const u32 c = c0 + tid;
                u32 cl = len > c ? (len - c + C - 1) / C : 0;
                if (ms) cl = len >> 1;
                cs.n = (int)cl;
                cs.msmode = ms ? (c == 0 ? 1 : 2) : 0;
                cs.sel_lo = cs.msmode == 2 ? 0x00ff : 0x0001;
                cs.sel_hi = cs.msmode == 2 ? 0xff00 : 0x0100;
                cs.pa = ms ? planes : planes + (size_t)c * stride;
                cs.pb = planes + stride;
                cs.glob = planes != smem_planes;
            }
            for (int i = tid; i < GROUP * (NWARP + 1); i += NT) (&s.edge[0][0])[i] = 0;
            __syncthreads();
            for (int q = 0; q < nchp; q++) {
                const ChanResult &rq = s.wres[q];
                chdr[q] = chan_hdr_bytes(all_raw, rq);
                cpos[q] = pos;
                pos += 4 + chdr[q] + rq.nbytes;
            }
            if (tid < nchp * MAXORD) {
                const ChanResult &rq = s.wres[tid / MAXORD];
                s.wqd[tid / MAXORD][tid % MAXORD] = ldexp((double)rq.coef[tid % MAXORD], -rq.shift);
            }
            __syncthreads();
            const ChanResult &r = s.wres[cq];
            if (wi == 0 && (tid & 31) == 0) {
                const u64 cp = cpos[cq];
                put_u32le(o + cp, chdr[cq] + r.nbytes);
                if (!all_raw) {
                    uint8_t *h = o + cp + 4;
                    const u32 nco = r.kind == 2 ? (u32)r.order : 0u;
                    *h++ = (uint8_t)nco;
                    for (u32 j = 0; j < nco; j++) { put_u32le(h, (u32)r.coef[j]); h += 4; }
                    *h++ = (uint8_t)(r.kind == 2 ? r.shift : (r.kind == 1 ? 128 + r.order : 0));   // encoder.rs:243, 279
                    *h++ = (uint8_t)(r.kind == 0 ? 2 : 0);                                         // ResidualEncoding
                    if (r.kind != 0) *h++ = (uint8_t)r.k;
                }
            }
            const u64 pay = cpos[cq] + 4 + chdr[cq];
            if (r.kind != 3 && r.nbytes > 0) pack_region<P>(s, ring + wid * WRING, s.cs[cq], cq, r, o, pay, wi, W, p.err);
            __syncthreads();
            // words shared by two regions: one lane per channel merges what the regions left in the edge accumulators
            if (wi == 0 && (tid & 31) == 0 && r.kind != 3 && r.nbytes > 0) {
                const u64 abase = pay & ~15ull, lo = pay, hi = pay + r.nbytes;
                u64 bp = (pay & 15ull) * 8ull;
                u32 curw = 0xffffffffu, curv = 0;
                auto flushw = [&]() {
                    if (curw != 0xffffffffu) {
                        const u64 a = abase + 4ull * curw;
                        for (int b = 0; b < 4; b++)
                            if (a + b >= lo && a + b < hi) o[a + b] = (uint8_t)(curv >> (24 - 8 * b));
                    }
                };
                for (int bd = 1; bd < W; bd++) {
                    bp += r.regbits[bd - 1];
                    if ((bp & 31) == 0) continue;
                    const u32 w = (u32)(bp >> 5);
                    if (w != curw) { flushw(); curw = w; curv = 0; }
                    curv |= s.edge[cq][bd];
                }
                flushw();
            }
        }
        if (tid == 0 && pos - fpos != fsize) atomicExch(p.err, 0xBAD00002u);
        __syncthreads();
        crc_frame(s, p, o, g, fpos, fsize);
        PH(const long long tc3b = clock64();)
        if (defer) {
            if (tid < 32) {
                u64 ex = lookback_wait(p.status, g, fsize);
                if (tid == 0) { s.frame_excl = ex; p.frame_excl[g] = ex; p.frame_size[g] = fsize; }
            }
            __syncthreads();
            PH(if (tid == 0) atomicAdd(p.phase_cycles + 15, (u64)(clock64() - tc3b));)
            copy_frame_out(o, p.out + data_base + s.frame_excl, fsize);
        }
        PH(if (tid == 0) {
            const long long tc4 = clock64();
            atomicAdd(p.phase_cycles + 0, (u64)(tc1 - tc0)); atomicAdd(p.phase_cycles + 1, (u64)(tc2 - tc1));
            atomicAdd(p.phase_cycles + 2, (u64)(tc3 - tc2)); atomicAdd(p.phase_cycles + 3, (u64)(tc4 - tc3));
            atomicAdd(p.phase_cycles + 4, (u64)(tc4 - tc0));
        })
    }
    __syncthreads();
    if (tid < 8 && s.cnt[tid]) atomicAdd(p.counters + tid, s.cnt[tid]);
}

// ----------------------------------------------------------------------------
// launch glue of this variant
// ----------------------------------------------------------------------------
#ifdef FLO_XBUILD_P8_ONLY       // experiment builds (tools/xbuild.py): levels 0-3 and 5-6 only
static cudaError_t variant_configure(size_t dyn_smem) {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(k_encode_frames<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem))) return e;
    return cudaFuncSetAttribute(k_encode_frames<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
}
static cudaError_t variant_launch(const EncodeParams &p, int grid, size_t dyn_smem, cudaStream_t st) {
    if (p.frame_end <= p.frame_begin) return cudaSuccess;
    if (p.level == 5 || p.level == 6) k_encode_frames<8><<<grid, NT, dyn_smem, st>>>(p);
    else if (p.level < 4) k_encode_frames<0><<<grid, NT, dyn_smem, st>>>(p);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}
#else
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
#endif
static int variant_occupancy(size_t dyn_smem) {
    int n = -1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_encode_frames<8>, NT, dyn_smem);
    return n;
}
