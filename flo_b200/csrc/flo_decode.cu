// flo_decode.cu -- lossless decoder kernels.
//
// Restates, for the GPU, what Reader::read (libflo/src/reader.rs:16-247) and Decoder::decode_file
// (libflo/src/lossless/decoder.rs:21-273) do on the CPU.  Rice decoding (core/rice.rs:123-159) and LPC synthesis
// (decoder.rs:152-184) are serial per channel, so the unit of parallel work is one channel of one frame:
// one lane per unit, 32 units per warp, all lanes stepping sample by sample in lock step.  That makes the two
// channels of a stereo frame neighbours in a warp, so the mid/side inverse is one shuffle and every lane writes
// its own channel of the interleaved f32 output directly -- no intermediate planes.
#include "flo_internal.h"

namespace flo {
namespace {

constexpr uint32_t FULL = 0xFFFFFFFFu;
constexpr uint8_t FT_TRANSFORM = 253, FT_RAW = 254;

__device__ __forceinline__ uint32_t rd32(const uint8_t *p) {
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}
__device__ __forceinline__ unsigned long long rd64(const uint8_t *p) { return (unsigned long long)rd32(p) | ((unsigned long long)rd32(p + 4) << 32); }
__device__ __forceinline__ void dec_fail(uint32_t *ctl, uint32_t frame, uint32_t chan1, uint32_t kind) {
    atomicMin(&ctl[1], (frame << 13) | (chan1 << 4) | kind);
}

// One thread per TOC entry: read_data_chunk / read_frame (reader.rs:101-166) up to the channel payloads.
__global__ void k_dec_parse(DecodeParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_toc) return;
    const uint8_t *f = p.file;
    const unsigned long long off = rd64(f + p.toc_pos + 20ull * i + 4);
    const unsigned long long fstart = p.data_start + off;
    DecFrame fr = {0u, 0u};
    if (fstart < p.data_start || fstart >= p.data_end) {          // reader.rs:116-118: stop at the first such entry
        atomicMin(&p.ctl[0], i);
        p.frames[i] = fr;
        return;
    }
    if (fstart + 6 > p.len) { dec_fail(p.ctl, i, 0, DEC_EOF); p.frames[i] = fr; return; }
    const uint32_t type = f[fstart], n = rd32(f + fstart + 1), flags = f[fstart + 5];
    fr.type_flags = type | (flags << 8);
    fr.n = n;
    p.frames[i] = fr;
    if (type == FT_TRANSFORM) { dec_fail(p.ctl, i, 0, DEC_TRANSFORM); return; }
    unsigned long long pos = fstart + 6;
    for (uint32_t c = 0; c < p.channels; c++) {
        DecUnit u = {0ull, 0u, 0u};
        if (pos + 4 > p.len) { dec_fail(p.ctl, i, c + 1, DEC_EOF); break; }
        const uint32_t ch_size = rd32(f + pos);
        const unsigned long long payload = pos + 4, ch_end = payload + ch_size;
        if (n > 2000000u) { dec_fail(p.ctl, i, c + 1, DEC_TOO_MANY); break; }       // reader.rs:175-177
        if (type >= 1 && type <= 12) {
            // the reader's own order of checks (reader.rs:208-214): the order byte is read and tested before anything
            // behind it is touched, and Reader::read meets every frame before the decoder runs -- so a bad order in an
            // earlier frame wins over a truncated later one, and over a truncated payload of the same channel
            if (payload + 1 > p.len) { dec_fail(p.ctl, i, c + 1, DEC_EOF); break; }
            if (f[payload] > 12) { dec_fail(p.ctl, i, c + 1, DEC_BAD_ORDER); break; }
            if (ch_end > p.len) { dec_fail(p.ctl, i, c + 1, DEC_EOF); break; }       // header or residual read runs off the file
        } else if (type == FT_RAW) {
            const unsigned long long need = 2ull * n;
            if (payload + (need < ch_size ? need : ch_size) > p.len) { dec_fail(p.ctl, i, c + 1, DEC_EOF); break; }
        }
        u.pos = payload; u.size = ch_size;
        p.units[(size_t)i * p.channels + c] = u;
        pos = ch_end;
    }
}

// Exclusive scan of frame_samples over the frames the reader keeps (one CTA; at most 100000 entries, reader.rs:86).
__global__ void __launch_bounds__(1024) k_dec_scan(DecodeParams p) {
    __shared__ unsigned long long wtot[32];
    __shared__ unsigned long long s_carry;
    const uint32_t tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const uint32_t keep = p.ctl[0] < p.n_toc ? p.ctl[0] : p.n_toc;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t b0 = 0; b0 < p.n_toc; b0 += 1024) {
        const uint32_t i = b0 + tid;
        unsigned long long v = 0;
        if (i < keep) v = p.frames[i].n;
        else if (i < p.n_toc) p.frames[i].n = 0;
        unsigned long long inc = v;
        #pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(FULL, inc, d);
            if ((int)lane >= d) inc += t;
        }
        if (lane == 31) wtot[w] = inc;
        __syncthreads();
        unsigned long long wb = 0, tot = 0;
        for (uint32_t k = 0; k < 32; k++) { unsigned long long t = wtot[k]; if (k < w) wb += t; tot += t; }
        const unsigned long long carry = s_carry;
        if (i < p.n_toc) p.base[i] = carry + wb + inc - v;
        __syncthreads();
        if (tid == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (tid == 0) *reinterpret_cast<unsigned long long *>(p.ctl + 2) = s_carry;
}

// ---- bit reader: MSB-first (rice.rs:217-260), bits past the end of the payload read as 0 ----
// Every lane reads its own stream, at its own pace.  A register written by a global load inside a divergent
// branch would stall the whole warp on the next lane's turn (the scoreboard is per warp register), so the
// streams are staged through shared memory instead: a 32-word ring per lane (word-interleaved, bank = lane),
// topped up at warp-uniform points with 16-byte loads that are stored to the ring one top-up later.  The
// per-sample path is branch-free: one predicated 32-bit append and one unconditional LDS of the next word.
constexpr int RING = 128;             // words per lane
constexpr int GROUP = 16;             // samples between top-ups (long enough for the loads of one top-up to land before the
                                      // next stores them); the fast path pops at most one word per sample
struct BitIn {
    const uint4 *gp, *vend;           // next vector to request; end of the readable image
    uint4 p0, p1, p2, p3;             // vectors in flight
    uint32_t npend;
    uint32_t wr, rd;                  // ring word counters: next word to store / word held in nxt
    unsigned long long end_off;       // payload end, as a byte offset from the aligned base of the stream
    uint32_t nxt;                     // ring[rd] (big-endian, ready to append), loaded ahead
    unsigned long long buf;           // MSB-aligned window
    int nb;                           // bits in the window
};
// The ring is addressed in the shared window directly (rs = this lane's column, 128 bytes between its words).
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ring_at(uint32_t rs, uint32_t idx) { return lds32(rs + ((idx & (RING - 1)) << 7)); }
// The ring holds big-endian words: the swap happens where a word is stored (once per word, a volatile PRMT right in front of
// the volatile store, so it cannot be hoisted up to the load), not on the per-code path of the reader.
__device__ __forceinline__ uint32_t be_store(uint32_t raw) { uint32_t v; asm volatile("prmt.b32 %0, %1, 0, 0x0123;" : "=r"(v) : "r"(raw)); return v; }
// Past the end of the image the address is clamped (the bytes are masked off in put_vec anyway): nothing may
// depend on the loaded registers until the next top-up stores them.
__device__ __forceinline__ uint4 ldv(const uint4 *p, const uint4 *end) { return __ldg(p < end ? p : end - 1); }
// Words go into the ring byte-swapped (big-endian), by a volatile PRMT directly in front of the store.  Nothing else here may touch the loaded
// registers except the stores themselves, or the compiler hoists that work to right behind the loads and the
// warp waits out the full memory latency at every top-up.
__device__ __forceinline__ uint32_t le_masked(uint32_t raw, unsigned long long pos, unsigned long long end_off) {
    if (pos + 4 <= end_off) return raw;
    return pos >= end_off ? 0u : (raw & ~(0xFFFFFFFFu << (8u * (uint32_t)(end_off - pos))));
}
__device__ __forceinline__ void put_vec(uint32_t rs, BitIn &b, const uint4 &v) {
    const uint32_t slot = rs + ((b.wr & (RING - 1)) << 7);  // wr is a multiple of 4: the four slots do not wrap
    const unsigned long long pos = 4ull * b.wr;
    if (pos + 16 <= b.end_off) {
        sts32(slot, be_store(v.x)); sts32(slot + 128, be_store(v.y)); sts32(slot + 256, be_store(v.z)); sts32(slot + 384, be_store(v.w));
    } else {                                               // the payload ends inside this vector: later bytes read as 0
        sts32(slot, be_store(le_masked(v.x, pos, b.end_off))); sts32(slot + 128, be_store(le_masked(v.y, pos + 4, b.end_off)));
        sts32(slot + 256, be_store(le_masked(v.z, pos + 8, b.end_off))); sts32(slot + 384, be_store(le_masked(v.w, pos + 12, b.end_off)));
    }
    b.wr += 4;
}
__device__ __forceinline__ void store_pending(uint32_t rs, BitIn &b) {
    if (b.npend >= 1) put_vec(rs, b, b.p0);
    if (b.npend >= 2) put_vec(rs, b, b.p1);
    if (b.npend >= 3) put_vec(rs, b, b.p2);
    if (b.npend >= 4) put_vec(rs, b, b.p3);
    b.npend = 0;
}
__device__ __forceinline__ void topup_once(uint32_t rs, BitIn &b) {
    store_pending(rs, b);
    const uint32_t room = RING - (b.wr - b.rd);
    if (room >= 4) { b.p0 = ldv(b.gp, b.vend); b.gp++; b.npend = 1; }
    if (room >= 8) { b.p1 = ldv(b.gp, b.vend); b.gp++; b.npend = 2; }
    if (room >= 12) { b.p2 = ldv(b.gp, b.vend); b.gp++; b.npend = 3; }
    if (room >= 16) { b.p3 = ldv(b.gp, b.vend); b.gp++; b.npend = 4; }
}
// Warp-uniform, once per TWO groups.  Afterwards every lane holds at least 4 * GROUP words, so the fast path (at most one word
// per code) cannot run dry before the next call.
__device__ __forceinline__ void topup(uint32_t rs, BitIn &b) {
    do topup_once(rs, b);
    while (__any_sync(FULL, b.wr - b.rd < 4 * GROUP));      // further trips (high bit rates): wait for the loads just issued
}
__device__ __forceinline__ void bits_init(uint32_t rs, BitIn &b, const uint8_t *file, unsigned long long len,
                                          unsigned long long start, uint32_t nbytes) {
    const uintptr_t a = (uintptr_t)(file + start), a0 = a & ~(uintptr_t)15;
    const uint32_t sk = (uint32_t)(a - a0);
    b.vend = (const uint4 *)(((uintptr_t)(file + len) + 15) & ~(uintptr_t)15);
    b.gp = nbytes ? (const uint4 *)a0 : b.vend;
    b.end_off = (unsigned long long)sk + nbytes;
    b.npend = 0; b.wr = 0; b.rd = 0;
    b.p0 = b.p1 = b.p2 = b.p3 = make_uint4(0u, 0u, 0u, 0u);
    topup(rs, b);
    b.rd = sk >> 2;                                        // the payload starts sk bytes into the first vector
    const uint32_t sb = sk & 3u;
    const uint32_t w = ring_at(rs, b.rd);
    b.rd++;
    b.buf = sb ? (unsigned long long)(w << (8u * sb)) << 32 : (unsigned long long)w << 32;
    b.nb = 32 - 8 * (int)sb;
    b.nxt = ring_at(rs, b.rd);
}
// lane-private (divergent) fill: at least `want` words on hand; fetches synchronously when the ring is dry
__device__ __forceinline__ void lane_fill(uint32_t rs, BitIn &b, int want) {
    while ((int)(b.wr - b.rd) < want) {
        if (b.npend) store_pending(rs, b);
        else { b.p0 = ldv(b.gp, b.vend); b.gp++; b.npend = 1; }
    }
}
__device__ __forceinline__ void refill_slow(uint32_t rs, BitIn &b) {
    if (b.nb > 32) return;
    b.buf |= (unsigned long long)b.nxt << (32 - b.nb); b.nb += 32; b.rd++;
    b.nxt = ring_at(rs, b.rd);
}
// decode_i32's loop body (rice.rs:127-155): unary quotient (ones, capped at 256 reads), k-bit remainder, zigzag.
// Rare path: the code does not fit the bits on hand (long unary run, large k).  One code takes at most
// 256 + 1 + 31 bits = 9 words; with 20 on hand at entry the rest of the group still pops without checking.
__device__ __forceinline__ uint32_t rice_slow(uint32_t rs, BitIn &b, uint32_t k) {
    lane_fill(rs, b, 2 * GROUP + 12);                     // this code and what is left of the two groups between top-ups
    b.nxt = ring_at(rs, b.rd);
    uint32_t q = 0;
    for (;;) {
        int run = __clzll((long long)~b.buf);
        if (run > b.nb) run = b.nb;
        const bool more = run == b.nb;                     // every loaded bit is a one: the run goes on
        if (q + (uint32_t)run >= 256u) {                   // rice.rs:139-141: stops after the 256th one, no terminator read
            const int take = (int)(256u - q);
            b.buf = take >= 64 ? 0ull : b.buf << take; b.nb -= take; q = 256u;
            break;
        }
        q += (uint32_t)run;
        if (!more) { b.buf = (b.buf << run) << 1; b.nb -= run + 1; break; }
        b.buf = 0; b.nb = 0;
        refill_slow(rs, b);
    }
    refill_slow(rs, b);
    uint32_t r = 0;
    if (k) { r = (uint32_t)(b.buf >> (64 - k)); b.buf <<= k; b.nb -= (int)k; }
    return (q << k) | r;
}
// Common path, free of branches: append the word on hand when the window is half empty (predicated), count the
// leading ones of the top word, cut the k remainder bits out with funnel shifts.  Returns the zigzag code
// (q << k | remainder).  When the code does not fit the bits on hand (ok false) nothing is consumed.
__device__ __forceinline__ uint32_t rice_try(uint32_t rs, BitIn &b, uint32_t k, bool &ok) {
    uint32_t hi = (uint32_t)(b.buf >> 32), lo = (uint32_t)b.buf;
    if (b.nb <= 32) {                                      // the low word of the window is empty
        const uint32_t w = b.nxt;
        hi |= __funnelshift_rc(w, 0u, (uint32_t)b.nb);
        lo = __funnelshift_lc(0u, w, 32u - (uint32_t)b.nb);
        b.nb += 32;
        b.rd++;
    }
    b.nxt = ring_at(rs, b.rd);                             // not needed before the next sample
    uint32_t run;                                          // leading ones of the top word; 0xFFFFFFFF when it is all ones
    asm("bfind.shiftamt.u32 %0, %1;" : "=r"(run) : "r"(~hi));
    const int used = (int)(run + 1u + k);
    ok = run < 32u && used <= b.nb;
    const uint32_t t = __funnelshift_lc(lo, hi, run + 1u);
    const uint32_t u = (run << k) | __funnelshift_rc(t, 0u, 32u - k);
    const unsigned long long w64 = ((unsigned long long)hi << 32) | lo;
    const int take = ok ? used : 0;                        // used <= 63 when ok
    b.buf = w64 << take;
    b.nb -= take;
    return u;
}
// The same step for the group transaction of the bit-reading warp, which rolls the reader back when any code of the
// group failed: nothing has to be preserved on failure, so the step is shorter.  After the append the window holds
// more than 32 bits, so a code of at most 32 bits (run <= 31 - k, one unsigned compare that also catches the all-ones
// word) always fits and lies entirely in the top word; window and bit count advance unconditionally (garbage after a
// failure, never an unsafe access: shifts clamp, the ring index is masked).
struct RiceK { uint32_t k, kp1, lim, pk; };               // k, k + 1, 31 - k, 2^k
__device__ __forceinline__ uint32_t rice_step(uint32_t rs, uint32_t &hi, uint32_t &lo, int &nb, uint32_t &rd, uint32_t &nxt,
                                              const RiceK &K, bool &all_ok) {
    // Append the look-ahead word when the low word of the window is empty: (w : 0) >> nb, upper word into hi, lower word is
    // the new lo.  The shift clamps at 32, so with more than 32 bits on hand the upper word is 0 and hi takes it
    // unconditionally -- the compare is off the chain that leads to the next leading-ones count.
    hi |= __funnelshift_rc(nxt, 0u, (uint32_t)nb);
    if (nb <= 32) {
        lo = __funnelshift_rc(0u, nxt, (uint32_t)nb);
        nb += 32;
        rd++;
    }
    nxt = ring_at(rs, rd);
    uint32_t run;
    asm("bfind.shiftamt.u32 %0, %1;" : "=r"(run) : "r"(~hi));
    all_ok = all_ok && run <= K.lim;
    const uint32_t used = run + K.kp1;
    const uint32_t t = __funnelshift_lc(0u, hi, run + 1u);                 // hi << (run + 1)
    const uint32_t u = run * K.pk + __funnelshift_rc(t, 0u, 32u - K.k);    // remainder < 2^k: add = or
    hi = __funnelshift_lc(lo, hi, used);
    lo = __funnelshift_lc(0u, lo, used);
    nb -= (int)used;
    return u;
}
__device__ __forceinline__ uint32_t rice_next_u(uint32_t rs, BitIn &b, uint32_t k) {
    bool ok;
    uint32_t u = rice_try(rs, b, k, ok);
    if (__builtin_expect(!ok, 0)) u = rice_slow(rs, b, k);
    return u;
}
__device__ __forceinline__ int32_t unzigzag(uint32_t u) { return (int32_t)(u >> 1) ^ -(int32_t)(u & 1u); }
__device__ __forceinline__ uint32_t zigzag(int32_t r) { return ((uint32_t)r << 1) ^ (uint32_t)(r >> 31); }
__device__ __forceinline__ int32_t rice_next(uint32_t rs, BitIn &b, uint32_t k) { return unzigzag(rice_next_u(rs, b, k)); }
// Four codes as one transaction: the common path runs without a branch; if any of the four did not fit, the
// reader state is rolled back and the four are redone one by one with the long-code path available.
__device__ __forceinline__ void rice_quad(uint32_t rs, BitIn &b, uint32_t k, uint32_t (&u)[4]) {
    const unsigned long long buf0 = b.buf;
    const int nb0 = b.nb;
    const uint32_t rd0 = b.rd, nxt0 = b.nxt;
    bool ok0, ok1, ok2, ok3;
    u[0] = rice_try(rs, b, k, ok0);
    u[1] = rice_try(rs, b, k, ok1);
    u[2] = rice_try(rs, b, k, ok2);
    u[3] = rice_try(rs, b, k, ok3);
    if (__builtin_expect(!(ok0 && ok1 && ok2 && ok3), 0)) {
        b.buf = buf0; b.nb = nb0; b.rd = rd0; b.nxt = nxt0;
        #pragma unroll 1
        for (int j = 0; j < 4; j++) {
            const uint32_t v = rice_next_u(rs, b, k);
            if (j == 0) u[0] = v; else if (j == 1) u[1] = v; else if (j == 2) u[2] = v; else u[3] = v;
        }
    }
}

__constant__ int c_fixed[5][4] = {{0, 0, 0, 0}, {1, 0, 0, 0}, {2, -1, 0, 0}, {3, -3, 1, 0}, {4, -6, 4, -1}};   // decoder.rs:199-259

enum { M_ZERO = 0, M_RICE = 1, M_PCM = 2 };

struct Lane {
    BitIn bits;
    uint32_t rs;                      // this lane's ring column (shared-window address)
    uint32_t msa, msb, mshift;        // output = (s * msa + neighbour * msb) / 2^mshift, truncating (mid/side inverse)
    const uint8_t *pcm; uint32_t pcm_bytes;
    uint8_t *outp; uint32_t stride;   // this lane's channel of the interleaved output (f32 or i16 elements), element stride
    uint32_t n, k;
    int src;                          // M_*
    int order, shift;                 // taps in use; >> shift (0 for the fixed predictors)
    bool fixed, ms, odd;
};

__device__ __forceinline__ int32_t next_residual(Lane &L, uint32_t i) {
    int32_t r = rice_next(L.rs, L.bits, L.k);
    if (L.src == M_PCM) {                                  // decoder.rs:132-143
        r = 0;
        if (2ull * i + 1 < L.pcm_bytes) r = (int16_t)((uint16_t)__ldg(L.pcm + 2ull * i) | ((uint16_t)__ldg(L.pcm + 2ull * i + 1) << 8));
    }
    return r;
}
// mid/side inverse (decoder.rs:75-89), i32 -> f32 (audio_constants.rs:24-26).
// `o` is the neighbour lane's sample (the other channel of a stereo frame): L = (m + s) / 2 on the even lane,
// R = (m - s) / 2 on the odd one, `/` truncating toward zero, all in wrapping 32-bit arithmetic; other frames
// pass s through.  One multiply-add form for all three so that the sample loop has no selects.
// OUT = float: the reference's output.  OUT = int16_t: the same integer sample before i32_to_f32, saturated to 16 bits
// (flo_decode_i16: half the bytes to store and to copy to the host).
template <typename OUT>
__device__ __forceinline__ OUT to_output(const Lane &L, int32_t s, int32_t o) {
    const uint32_t t = (uint32_t)s * L.msa + (uint32_t)o * L.msb;
    const int32_t v = (int32_t)(t + ((t >> 31) & L.mshift)) >> L.mshift;
    if constexpr (sizeof(OUT) == 2) return (OUT)max(-32768, min(32767, v));
    else return __fmul_rn(__int2float_rn(v), 1.0f / 32767.0f);
}
__device__ __forceinline__ void store_if(float *p, float v, bool on) {     // predicated store: no branch in the sample loop
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %2, 0;\n @q st.global.f32 [%0], %1;\n}" :: "l"(p), "f"(v), "r"((uint32_t)on) : "memory");
}
__device__ __forceinline__ void store_if(int16_t *p, int16_t v, bool on) {
    asm volatile("{\n .reg .pred q;\n setp.ne.u32 q, %2, 0;\n @q st.global.b16 [%0], %1;\n}" :: "l"(p), "h"(v), "r"((uint32_t)on) : "memory");
}
// Store v at p + T * step bytes if T < rem: the compare lives inside the asm (one SETP, no boolean materialised), the
// address is one multiply-add from the block's base.
template <int T>
__device__ __forceinline__ void store_lt(uint8_t *base, uint32_t step_bytes, float v, uint32_t rem) {
    float *p = reinterpret_cast<float *>(base + (size_t)T * step_bytes);
    asm volatile("{\n .reg .pred q;\n setp.gt.u32 q, %2, %3;\n @q st.global.f32 [%0], %1;\n}" :: "l"(p), "f"(v), "r"(rem), "n"(T) : "memory");
}
template <int T>
__device__ __forceinline__ void store_lt(uint8_t *base, uint32_t step_bytes, int16_t v, uint32_t rem) {
    int16_t *p = reinterpret_cast<int16_t *>(base + (size_t)T * step_bytes);
    asm volatile("{\n .reg .pred q;\n setp.gt.u32 q, %2, %3;\n @q st.global.b16 [%0], %1;\n}" :: "l"(p), "h"(v), "r"(rem), "n"(T) : "memory");
}
template <typename OUT>
__device__ __forceinline__ void emit(const Lane &L, uint32_t i, int32_t s) {
    store_if(reinterpret_cast<OUT *>(L.outp) + (size_t)i * L.stride, to_output<OUT>(L, s, __shfl_xor_sync(FULL, s, 1)), i < L.n);
}

// ---- predictor side ----
constexpr int BLK = GROUP;            // samples handed from the bit-reading warp to the predictor warp at a time

template <int ORD>
__device__ __forceinline__ int32_t fir(const int32_t (&c)[ORD], const int32_t (&h)[ORD], int sh, int32_t r) {
    long long a0 = 0, a1 = 0;                              // two chains; the newest sample enters last
    #pragma unroll
    for (int j = ORD - 1; j >= 0; j--) {
        if (j & 1) a1 += (long long)c[j] * (long long)h[j];
        else a0 += (long long)c[j] * (long long)h[j];
    }
    return (int32_t)((uint32_t)(int32_t)((a0 + a1) >> sh) + (uint32_t)r);
}
// Steady state (sample index >= 12 >= order): reconstruct_lpc_int's loop (decoder.rs:169-179) / the fixed
// recurrences (decoder.rs:199-259) as one FIR of at most ORD taps.  Fully unrolled over the block, so the history
// is renamed instead of moved and the shuffle / convert / store of sample t overlap the filter step of t + 1.
template <int ORD, typename OUT, int T>
__device__ __forceinline__ void consume_one(const Lane &L, uint32_t res, const int32_t (&c)[ORD], int32_t (&h)[ORD], uint8_t *base, uint32_t step_bytes, uint32_t rem) {
    const int32_t s = fir<ORD>(c, h, L.shift, unzigzag(lds32(res + 128u * T)));
    #pragma unroll
    for (int j = ORD - 1; j > 0; j--) h[j] = h[j - 1];
    h[0] = s;
    store_lt<T>(base, step_bytes, to_output<OUT>(L, s, __shfl_xor_sync(FULL, s, 1)), rem);
    if constexpr (T + 1 < BLK) consume_one<ORD, OUT, T + 1>(L, res, c, h, base, step_bytes, rem);
}
template <int ORD, typename OUT>
__device__ __forceinline__ void consume_block(const Lane &L, uint32_t res, uint32_t i0, const int32_t (&c)[ORD], int32_t (&h)[ORD], OUT *&op) {
    const uint32_t rem = L.n > i0 ? L.n - i0 : 0u, step_bytes = L.stride * (uint32_t)sizeof(OUT);
    consume_one<ORD, OUT, 0>(L, res, c, h, reinterpret_cast<uint8_t *>(op), step_bytes, rem);
    op += (size_t)BLK * L.stride;
}
// Generic step with the history newest-first in hist[]: warm-up rules (decoder.rs:163-165, 199-259).
template <typename OUT>
__device__ __forceinline__ void synth_apply(const Lane &L, uint32_t i, int32_t r, const int32_t (&c12)[12], int32_t (&hist)[12]) {
    int32_t pred = 0;
    if (i >= (uint32_t)L.order) {
        long long acc = 0;
        #pragma unroll
        for (int j = 0; j < 12; j++) acc += (long long)c12[j] * (long long)hist[j];
        pred = (int32_t)(acc >> L.shift);
    } else if (L.fixed) {                                  // sample i < order uses the order-i predictor
        long long acc = 0;
        #pragma unroll
        for (int j = 0; j < 4; j++) acc += (long long)c_fixed[i][j] * (long long)hist[j];
        pred = (int32_t)acc;
    }
    const int32_t s = (int32_t)((uint32_t)pred + (uint32_t)r);
    #pragma unroll
    for (int j = 11; j > 0; j--) hist[j] = hist[j - 1];
    hist[0] = s;
    emit<OUT>(L, i, s);
}
// Predictor warp: block b of residuals is consumed while the bit-reading warp produces block b + 1.
template <int ORD, typename OUT>
__device__ __forceinline__ void consumer_loop(const Lane &L, uint32_t res0, uint32_t nblk, const int32_t (&c12)[12], int32_t (&hist)[12]) {
    int32_t c[ORD], h[ORD];
    #pragma unroll
    for (int j = 0; j < ORD; j++) { c[j] = c12[j]; h[j] = 0; }
    OUT *op = reinterpret_cast<OUT *>(L.outp) + (size_t)BLK * L.stride;
    for (uint32_t ph = 0; ph <= nblk; ph++) {
        if (ph == 1) {                                     // first block: warm-up rules, generic taps
            #pragma unroll 1
            for (int t = 0; t < BLK; t++) synth_apply<OUT>(L, (uint32_t)t, unzigzag(lds32(res0 + 128u * t)), c12, hist);
            #pragma unroll
            for (int j = 0; j < ORD; j++) h[j] = hist[j];
        } else if (ph > 1) {
            consume_block<ORD, OUT>(L, res0 + (((ph - 1) & 1u) ? 128u * BLK : 0u), (ph - 1) * BLK, c, h, op);
        }
        __syncthreads();
    }
}
// Bit-reading warp: residual i of every lane, as its zigzag code -> res[block parity][i % BLK][lane].
__device__ __forceinline__ void producer_loop(Lane &L, uint32_t res0, uint32_t nblk, bool any_pcm) {
    const RiceK K = {L.k, L.k + 1u, 31u - L.k, 1u << L.k};           // k <= 31 (lane_setup)
    for (uint32_t ph = 0; ph <= nblk; ph++) {
        if (ph < nblk) {
            const uint32_t res = res0 + ((ph & 1u) ? 128u * BLK : 0u);
            #pragma unroll 1
            for (int g = 0; g < BLK; g += GROUP) {
                if ((ph & 1u) == 0) topup(L.rs, L.bits);           // every other group: half the top-up instructions per code
                if (!any_pcm) {
                    // The whole group as one transaction: GROUP codes straight-line with predication only; if any of
                    // them did not fit the bits on hand, roll the reader back and redo the group code by code.
                    BitIn &b = L.bits;
                    const unsigned long long buf0 = b.buf;
                    const int nb0 = b.nb;
                    const uint32_t rd0 = b.rd, nxt0 = b.nxt;
                    uint32_t u[GROUP];
                    bool all_ok = true;
                    uint32_t hi = (uint32_t)(buf0 >> 32), lo = (uint32_t)buf0, rd = rd0, nxt = nxt0;
                    int nb = nb0;
                    #pragma unroll
                    for (int t = 0; t < GROUP; t++) u[t] = rice_step(L.rs, hi, lo, nb, rd, nxt, K, all_ok);
                    if (__builtin_expect(!all_ok, 0)) {
                        b.buf = buf0; b.nb = nb0; b.rd = rd0; b.nxt = nxt0;
                        #pragma unroll 1
                        for (int t = 0; t < GROUP; t++) sts32(res + 128u * (g + t), rice_next_u(L.rs, b, L.k));
                    } else {
                        b.buf = ((unsigned long long)hi << 32) | lo; b.nb = nb; b.rd = rd; b.nxt = nxt;
                        #pragma unroll
                        for (int t = 0; t < GROUP; t++) sts32(res + 128u * (g + t), u[t]);
                    }
                } else {                                   // raw PCM lanes (rare): the generic source for the whole warp
                    #pragma unroll 1
                    for (int t = 0; t < GROUP; t++) sts32(res + 128u * (g + t), zigzag(next_residual(L, ph * BLK + g + t)));
                }
            }
        }
        __syncthreads();
    }
}

// One channel of one frame: read_channel_data's ALPC arm (reader.rs:207-244) and decode_channel_int's case
// split (decoder.rs:92-148) -> where the residual bits are, which predictor runs over them, where the samples go.
__device__ __forceinline__ void lane_setup(const DecodeParams &p, unsigned long long u, unsigned long long n_units, uint32_t lane,
                                           Lane &L, int32_t (&c12)[12], int32_t (&hist)[12], unsigned long long &rpos, uint32_t &rbytes) {
    const uint32_t C = p.channels;
    const uint8_t *f = p.file;
    L.n = 0; L.k = 0; L.src = M_ZERO; L.order = 0; L.shift = 0; L.fixed = false; L.ms = false; L.odd = (lane & 1u) != 0;
    L.pcm = f; L.pcm_bytes = 0; L.outp = (uint8_t *)p.out; L.stride = C;
    #pragma unroll
    for (int j = 0; j < 12; j++) { c12[j] = 0; hist[j] = 0; }
    rpos = 0; rbytes = 0;

    // an error in a frame the reader never reaches (behind its stop entry) does not count (reader.rs:116-118)
    const uint32_t errkey = p.ctl[1];
    const bool failed = errkey != 0xFFFFFFFFu && (errkey >> 13) < (p.ctl[0] < p.n_toc ? p.ctl[0] : p.n_toc);
    if (u < n_units && !failed) {
        const uint32_t fi = (uint32_t)(u / C), ch = (uint32_t)(u % C);
        const DecFrame fr = p.frames[fi];
        const DecUnit un = p.units[u];
        const uint32_t type = fr.type_flags & 0xFFu, flags = (fr.type_flags >> 8) & 0xFFu;
        L.n = fr.n;
        L.ms = C == 2 && (flags & 1u);
        L.outp = (uint8_t *)p.out + ((size_t)p.base[fi] * C + ch) * (p.out_i16 ? 2u : 4u);
        const unsigned long long ch_end = un.pos + un.size;
        if (type == FT_RAW) {                              // reader.rs:182-188
            const unsigned long long need = 2ull * fr.n;
            L.pcm_bytes = (uint32_t)(need < un.size ? need : un.size);
            L.pcm = f + un.pos;
            if (L.pcm_bytes) L.src = M_PCM;
        } else if (type >= 1 && type <= 12) {              // reader.rs:207-244
            unsigned long long pos = un.pos;
            bool ok = true;
            auto eof = [&](unsigned long long need_end) { if (need_end > p.len) { dec_fail(p.ctl, fi, ch + 1, DEC_EOF); ok = false; } return !ok; };
            uint32_t order = 0, ncoef = 0, shift = 0, eb = 0, k = 0;
            if (!eof(pos + 1)) {
                order = f[pos++];
                if (order > 12) { dec_fail(p.ctl, fi, ch + 1, DEC_BAD_ORDER); ok = false; }
            }
            if (ok) {                                      // coefficients stop at the channel end (reader.rs:218-223)
                const unsigned long long fit = ch_end > pos ? (ch_end - pos) / 4 : 0;
                ncoef = fit < order ? (uint32_t)fit : order;
                if (!eof(pos + 4ull * ncoef)) {
                    #pragma unroll
                    for (int j = 0; j < 12; j++) if ((uint32_t)j < ncoef) c12[j] = (int32_t)rd32(f + pos + 4 * j);
                    pos += 4ull * ncoef;
                }
            }
            if (ok && !eof(pos + 1)) shift = f[pos++];
            if (ok && !eof(pos + 1)) eb = f[pos++];
            if (ok && eb == 0 && !eof(pos + 1)) k = f[pos++];        // rice parameter only for ResidualEncoding::Rice
            const unsigned long long rem = ok && ch_end > pos ? ch_end - pos : 0;
            if (ok && rem) eof(pos + rem);
            if (ok) {                                      // decode_channel_int (decoder.rs:92-148)
                if (ncoef == 0 && rem && shift >= 128) {
                    const uint32_t fo = shift - 128;
                    L.src = M_RICE;
                    if (fo >= 1 && fo <= 4) {              // other orders copy the residuals (decoder.rs:195-198, 261-264)
                        L.fixed = true; L.order = (int)fo;
                        #pragma unroll
                        for (int j = 0; j < 4; j++) c12[j] = c_fixed[fo][j];
                    }
                } else if (ncoef) {
                    L.src = M_RICE; L.order = (int)ncoef; L.shift = (int)(shift & 63u);
                } else if (rem) {                          // raw PCM inside an ALPC frame (decoder.rs:132-144)
                    L.src = M_PCM; L.pcm = f + pos;
                    const unsigned long long need = 2ull * fr.n;
                    L.pcm_bytes = (uint32_t)(need < rem ? need : rem);
                }
                if (L.src == M_RICE) {
                    rpos = pos; rbytes = (uint32_t)rem; L.k = rem ? k : 0;     // no bytes: every residual is 0 (rice.rs:128-131)
                    if (L.k > 31) { dec_fail(p.ctl, fi, ch + 1, DEC_BAD_K); ok = false; }
                }
            }
            if (!ok) {
                L.src = M_ZERO; L.order = 0; L.shift = 0; L.fixed = false; L.k = 0;
                #pragma unroll
                for (int j = 0; j < 12; j++) c12[j] = 0;
            }
        }
        // Silence / reserved types: zeros (reader.rs:180, 246)
    }
    L.msa = L.ms ? (L.odd ? 0xFFFFFFFFu : 1u) : 1u;
    L.msb = L.ms ? 1u : 0u;
    L.mshift = L.ms ? 1u : 0u;
}

// One CTA = 32 units and two warps.  Rice decoding and the predictor are two serial chains per channel; run by one
// warp they add up (and a lone warp has nobody to hide its stalls behind).  Warp 0 runs the bit reader, warp 1 the
// predictor + mid/side + output, one block of BLK residuals apart, handing over through shared memory.
template <typename OUT>
__global__ void __launch_bounds__(64) k_dec_units(DecodeParams p) {
    __shared__ uint32_t ring[RING * 32];
    __shared__ uint32_t resbuf[2 * BLK * 32];
    const uint32_t lane = threadIdx.x & 31u;
    const bool reader = threadIdx.x < 32;
    const unsigned long long u = (unsigned long long)blockIdx.x * 32 + lane;
    const uint32_t keep = p.ctl[0] < p.n_toc ? p.ctl[0] : p.n_toc;
    const unsigned long long n_units = (unsigned long long)keep * p.channels;

    Lane L;
    int32_t c12[12], hist[12];
    unsigned long long rpos; uint32_t rbytes;
    lane_setup(p, u, n_units, lane, L, c12, hist, rpos, rbytes);
    const bool any_pcm = __any_sync(FULL, L.src == M_PCM);
    const uint32_t nmax = __reduce_max_sync(FULL, L.n);
    const int omax = (int)__reduce_max_sync(FULL, (uint32_t)L.order);
    const uint32_t nblk = (nmax + BLK - 1) / BLK;
    const uint32_t res0 = (uint32_t)__cvta_generic_to_shared(resbuf) + 4u * lane;
    if (reader) {
        L.rs = (uint32_t)__cvta_generic_to_shared(ring) + 4u * lane;
        bits_init(L.rs, L.bits, p.file, p.len, L.src == M_RICE ? rpos : 0ull, L.src == M_RICE ? rbytes : 0u);
        producer_loop(L, res0, nblk, any_pcm);
    } else {
        if (omax <= 4) consumer_loop<4, OUT>(L, res0, nblk, c12, hist);
        else if (omax <= 8) consumer_loop<8, OUT>(L, res0, nblk, c12, hist);
        else consumer_loop<12, OUT>(L, res0, nblk, c12, hist);
    }
}

}  // namespace

cudaError_t launch_decode_parse(const DecodeParams &p, cudaStream_t st) {
    if (p.n_toc) k_dec_parse<<<(p.n_toc + 127) / 128, 128, 0, st>>>(p);
    k_dec_scan<<<1, 1024, 0, st>>>(p);
    return cudaGetLastError();
}
cudaError_t launch_decode_units(const DecodeParams &p, cudaStream_t st) {
    const unsigned long long units = (unsigned long long)p.n_toc * p.channels;
    if (units) {
        if (p.out_i16) k_dec_units<int16_t><<<(unsigned)((units + 31) / 32), 64, 0, st>>>(p);
        else k_dec_units<float><<<(unsigned)((units + 31) / 32), 64, 0, st>>>(p);
    }
    return cudaGetLastError();
}

}  // namespace flo
