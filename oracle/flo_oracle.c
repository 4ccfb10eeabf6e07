/*
 * flo_oracle.c -- CPU oracle: a plain-C restatement of flo's lossless ALPC
 * encode path (and the decoder/reader needed to check decoded samples).
 *
 * TEST INFRASTRUCTURE ONLY (see flo_oracle.h).  Deliberately literal: the same
 * exhaustive candidate loop and the same bit-at-a-time Rice writer as the
 * reference, so that timing it is a fair "port" CPU baseline.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math (no FMA contraction: the
 * Levinson-Durbin recursion must round every product and every sum).
 *
 * Citations are reference file:line (relative to the reference checkout).
 */
#include "flo_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread char g_err[256];
static void set_err(const char *m) { snprintf(g_err, sizeof g_err, "%s", m); }
const char *flo_ref_last_error(void) { return g_err; }
void flo_ref_free(void *p) { free(p); }

/* ---------------------------------------------------------------- bytes -- */
typedef struct { uint8_t *p; size_t len, cap; } bytes_t;

static void bytes_reserve(bytes_t *b, size_t extra) {
    if (b->len + extra <= b->cap) return;
    size_t nc = b->cap ? b->cap * 2 : 64;
    while (nc < b->len + extra) nc *= 2;
    b->p = (uint8_t *)realloc(b->p, nc);
    b->cap = nc;
}
static void bytes_push(bytes_t *b, uint8_t v) { bytes_reserve(b, 1); b->p[b->len++] = v; }
static void bytes_extend(bytes_t *b, const void *src, size_t n) {
    if (!n) return;
    bytes_reserve(b, n); memcpy(b->p + b->len, src, n); b->len += n;
}
static void bytes_u16(bytes_t *b, uint16_t v) { uint8_t t[2] = { (uint8_t)v, (uint8_t)(v >> 8) }; bytes_extend(b, t, 2); }
static void bytes_u32(bytes_t *b, uint32_t v) { uint8_t t[4]; for (int i = 0; i < 4; i++) t[i] = (uint8_t)(v >> (8 * i)); bytes_extend(b, t, 4); }
static void bytes_u64(bytes_t *b, uint64_t v) { uint8_t t[8]; for (int i = 0; i < 8; i++) t[i] = (uint8_t)(v >> (8 * i)); bytes_extend(b, t, 8); }

/* ------------------------------------------------ core/audio_constants.rs -- */
/* f32_to_i32, audio_constants.rs:18-20:
 *   (sample * 32767.0f32).clamp(-32768.0, 32767.0) as i32
 * Rust f32::clamp keeps NaN; `as i32` truncates toward zero, NaN -> 0. */
int32_t flo_ref_f32_to_i32(float x) {
    volatile float y = x * 32767.0f;          /* one RN f32 multiply */
    float c = y;
    if (c != c) return 0;                      /* NaN survives clamp, `as i32` gives 0 */
    if (c < -32768.0f) c = -32768.0f;
    if (c > 32767.0f) c = 32767.0f;
    return (int32_t)c;                         /* C float->int conversion truncates */
}
/* i32_to_f32, audio_constants.rs:24-26: sample as f32 * (1.0/32767.0) */
float flo_ref_i32_to_f32(int32_t v) {
    const float scale = 1.0f / 32767.0f;
    return (float)v * scale;
}

/* --------------------------------------------------------- core/crc32.rs -- */
static uint32_t g_crc_table[256];
static int g_crc_ready;
static void crc_init(void) {                   /* crc32.rs:2-20 */
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i;
        for (int j = 0; j < 8; j++) c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1;
        g_crc_table[i] = c;
    }
    g_crc_ready = 1;
}
uint32_t flo_ref_crc32(const uint8_t *data, size_t len) { /* crc32.rs:23-30 */
    if (!g_crc_ready) crc_init();
    uint32_t crc = 0xFFFFFFFFu;
    for (size_t i = 0; i < len; i++) crc = (crc >> 8) ^ g_crc_table[(crc ^ data[i]) & 0xFF];
    return ~crc;
}

/* ---------------------------------------------------------- core/rice.rs -- */
static uint32_t unsigned_abs32(int32_t r) { return r < 0 ? 0u - (uint32_t)r : (uint32_t)r; }
static int bitlen64(uint64_t v) { int n = 0; while (v) { n++; v >>= 1; } return n; }

uint8_t flo_ref_estimate_rice_parameter_i32(const int32_t *r, size_t n) { /* rice.rs:29-69 */
    if (n == 0) return 4;
    uint64_t max_abs = 0;
    for (size_t i = 0; i < n; i++) { uint64_t a = unsigned_abs32(r[i]); if (a > max_abs) max_abs = a; }
    if (max_abs == 0) return 0;
    uint64_t max_unsigned = 2 * max_abs;
    int min_k = 0;
    if (max_unsigned > 255) {
        int bits_needed = bitlen64(max_unsigned);          /* 64 - leading_zeros */
        min_k = bits_needed > 8 ? bits_needed - 8 : 0;     /* saturating_sub(8) */
    }
    uint64_t sum = 0;
    for (size_t i = 0; i < n; i++) sum += unsigned_abs32(r[i]);
    uint32_t mean = (uint32_t)(sum / (uint64_t)n);
    int mean_k = mean > 0 ? bitlen64(mean) : 0;            /* 32 - leading_zeros */
    int k = min_k > mean_k ? min_k : mean_k;
    if (k > 15) k = 15;
    return (uint8_t)k;
}

/* BitWriter, rice.rs:162-208 (MSB-first, final partial byte zero padded) */
typedef struct { bytes_t bytes; uint8_t cur; uint8_t bit_pos; } bitw_t;
static void bw_bit(bitw_t *w, uint32_t bit) {
    if (bit) w->cur |= (uint8_t)(1u << (7 - w->bit_pos));
    if (++w->bit_pos == 8) { bytes_push(&w->bytes, w->cur); w->cur = 0; w->bit_pos = 0; }
}
static void bw_finish(bitw_t *w) { if (w->bit_pos > 0) bytes_push(&w->bytes, w->cur); }

/* encode_sample, rice.rs:94-114 */
static void rice_encode_sample(bitw_t *w, int32_t sample, uint8_t k) {
    uint32_t u = ((uint32_t)sample << 1) ^ (uint32_t)(sample >> 31);
    uint32_t q = u >> k;
    uint32_t rem = u & ((1u << k) - 1u);
    uint32_t qc = q < 255 ? q : 255;
    for (uint32_t i = 0; i < qc; i++) bw_bit(w, 1);
    bw_bit(w, 0);
    for (int i = (int)k - 1; i >= 0; i--) bw_bit(w, (rem >> i) & 1);
}
static bytes_t rice_encode(const int32_t *r, size_t n, uint8_t k) { /* rice.rs:84-92 */
    bitw_t w; memset(&w, 0, sizeof w);
    for (size_t i = 0; i < n; i++) rice_encode_sample(&w, r[i], k);
    bw_finish(&w);
    return w.bytes;
}
size_t flo_ref_rice_encode_i32(const int32_t *r, size_t n, uint8_t k, uint8_t **out) {
    bytes_t b = rice_encode(r, n, k);
    *out = b.p;
    return b.len;
}

/* BitReader + decode_i32, rice.rs:123-159, 217-269 */
typedef struct { const uint8_t *p; size_t len, byte_pos; uint8_t bit_pos; } bitr_t;
static uint32_t br_bit(bitr_t *r) {
    if (r->byte_pos >= r->len) return 0;
    uint32_t bit = (r->p[r->byte_pos] >> (7 - r->bit_pos)) & 1;
    if (++r->bit_pos == 8) { r->bit_pos = 0; r->byte_pos++; }
    return bit;
}
void flo_ref_rice_decode_i32(const uint8_t *enc, size_t enc_len, uint8_t k, size_t target_len, int32_t *out) {
    bitr_t r = { enc, enc_len, 0, 0 };
    for (size_t i = 0; i < target_len; i++) {
        if (r.byte_pos >= r.len) { out[i] = 0; continue; }
        uint32_t q = 0;
        while (r.byte_pos < r.len && br_bit(&r) == 1) { q++; if (q > 255) break; }
        uint32_t rem = 0;
        for (int j = 0; j < k; j++) rem = (rem << 1) | br_bit(&r);
        uint32_t u = (k < 32 ? (q << k) : 0) | rem;
        out[i] = (int32_t)(u >> 1) ^ -(int32_t)(u & 1);
    }
}

/* ------------------------------------------------------- lossless/lpc.rs -- */
void flo_ref_autocorr_int(const int32_t *s, size_t n, int order, int64_t *out) { /* lpc.rs:213-221 */
    for (int lag = 0; lag <= order; lag++) {
        int64_t acc = 0;
        for (size_t i = (size_t)lag; i < n; i++) acc += (int64_t)s[i] * (int64_t)s[i - lag];
        out[lag] = acc;
    }
}

/* Rust `f64 as u8`: saturating, NaN -> 0 */
static uint8_t f64_as_u8(double v) { if (v != v) return 0; if (v <= 0.0) return 0; if (v >= 255.0) return 255; return (uint8_t)v; }
/* Rust `f64 as i32`: saturating, NaN -> 0 */
static int32_t f64_as_i32(double v) {
    if (v != v) return 0;
    if (v <= -2147483648.0) return INT32_MIN;
    if (v >= 2147483647.0) return INT32_MAX;
    return (int32_t)v;
}

int flo_ref_levinson_durbin_int(const int64_t *ac, int order, int32_t *coeffs_fp, uint8_t *shift_out) { /* lpc.rs:225-276 */
    if (ac[0] == 0) return 0;
    double coeffs[32]; double nc[32];
    for (int i = 0; i < order; i++) coeffs[i] = 0.0;
    double error = (double)ac[0];
    for (int i = 0; i < order; i++) {
        double lambda = (double)ac[i + 1];
        for (int j = 0; j < i; j++) lambda -= coeffs[j] * (double)ac[i - j];
        if (fabs(error) < 1e-10) return 0;
        double gamma = lambda / error;
        if (fabs(gamma) >= 1.0) return 0;
        nc[i] = gamma;
        for (int j = 0; j < i; j++) nc[j] = coeffs[j] - gamma * coeffs[i - 1 - j];
        for (int j = 0; j <= i; j++) coeffs[j] = nc[j];
        error *= 1.0 - gamma * gamma;
    }
    double max_coeff = 0.0;
    for (int i = 0; i < order; i++) {           /* fold(0.0, f64::max): NaN operands are dropped by f64::max */
        double a = fabs(coeffs[i]);
        if (a != a) continue;
        if (a > max_coeff) max_coeff = a;
    }
    if (max_coeff == 0.0 || !isfinite(max_coeff)) return 0;
    uint8_t shift = f64_as_u8(floor(log2((double)(1 << 30) / max_coeff)));
    if (shift > 15) shift = 15;
    double scale = (double)((int64_t)1 << shift);
    for (int i = 0; i < order; i++) coeffs_fp[i] = f64_as_i32(round(coeffs[i] * scale)); /* f64::round: half away from zero */
    *shift_out = shift;
    return 1;
}

void flo_ref_calc_residuals_int(const int32_t *s, size_t n, const int32_t *coeffs, uint8_t shift, int order, int32_t *out) { /* lpc.rs:279-298 */
    size_t warm = (size_t)order < n ? (size_t)order : n;
    for (size_t i = 0; i < warm; i++) out[i] = s[i];
    for (size_t i = (size_t)order; i < n; i++) {
        int64_t pred = 0;
        for (int j = 0; j < order; j++) pred += (int64_t)coeffs[j] * (int64_t)s[i - (size_t)j - 1];
        pred >>= shift;                                        /* arithmetic shift */
        out[i] = (int32_t)((uint32_t)s[i] - (uint32_t)(int32_t)pred); /* `as i32` truncates; release-mode wrapping sub */
    }
}

/* wrapping helpers for the i32 expressions in fixed_predictor_residuals */
#define W(x) ((uint32_t)(x))
void flo_ref_fixed_predictor_residuals(const int32_t *s, size_t n, int order, int32_t *out) { /* lpc.rs:301-359 */
    if (n == 0) return;
    switch (order) {
    case 1:
        out[0] = s[0];
        for (size_t i = 1; i < n; i++) out[i] = (int32_t)(W(s[i]) - W(s[i - 1]));
        break;
    case 2:
        out[0] = s[0];
        if (n > 1) out[1] = (int32_t)(W(s[1]) - W(s[0]));
        for (size_t i = 2; i < n; i++) out[i] = (int32_t)(W(s[i]) - 2u * W(s[i - 1]) + W(s[i - 2]));
        break;
    case 3:
        out[0] = s[0];
        if (n > 1) out[1] = (int32_t)(W(s[1]) - W(s[0]));
        if (n > 2) out[2] = (int32_t)(W(s[2]) - 2u * W(s[1]) + W(s[0]));
        for (size_t i = 3; i < n; i++) out[i] = (int32_t)(W(s[i]) - 3u * W(s[i - 1]) + 3u * W(s[i - 2]) - W(s[i - 3]));
        break;
    case 4:
        out[0] = s[0];
        if (n > 1) out[1] = (int32_t)(W(s[1]) - W(s[0]));
        if (n > 2) out[2] = (int32_t)(W(s[2]) - 2u * W(s[1]) + W(s[0]));
        if (n > 3) out[3] = (int32_t)(W(s[3]) - 3u * W(s[2]) + 3u * W(s[1]) - W(s[0]));
        for (size_t i = 4; i < n; i++)
            out[i] = (int32_t)(W(s[i]) - 4u * W(s[i - 1]) + 6u * W(s[i - 2]) - 4u * W(s[i - 3]) + W(s[i - 4]));
        break;
    default:                                   /* order 0 and anything else: identity */
        memcpy(out, s, n * sizeof(int32_t));
        break;
    }
}

/* ----------------------------------------------------------- core/types.rs -- */
enum { FT_SILENCE = 0, FT_TRANSFORM = 253, FT_RAW = 254, FT_RESERVED = 255 };
enum { RE_RICE = 0, RE_GOLOMB = 1, RE_RAW = 2 };

typedef struct {
    int32_t coeffs[12]; uint32_t n_coeffs;
    uint8_t shift_bits, encoding, rice_parameter;
    bytes_t residuals;
} chan_t;

typedef struct {
    uint8_t frame_type; uint32_t frame_samples; uint8_t flags;
    uint32_t n_channels; chan_t *ch;
} frame_t;

static int ft_is_alpc(uint8_t t) { return t >= 1 && t <= 12; }
static uint8_t ft_from_order(size_t order) { return (order >= 1 && order <= 12) ? (uint8_t)order : 8; } /* types.rs:69-85 */

static void frame_free(frame_t *f) {
    for (uint32_t c = 0; c < f->n_channels; c++) free(f->ch[c].residuals.p);
    free(f->ch); f->ch = NULL; f->n_channels = 0;
}

static size_t frame_byte_size(const frame_t *f) {           /* types.rs:242-267 */
    size_t size = 6;
    for (uint32_t c = 0; c < f->n_channels; c++) {
        const chan_t *ch = &f->ch[c];
        size += 4;
        if (f->frame_type == FT_TRANSFORM) size += ch->residuals.len;
        else if (ft_is_alpc(f->frame_type)) {
            size += 1 + ch->n_coeffs * 4 + 1 + 1;
            if (ch->encoding == RE_RICE) size += 1;
            size += ch->residuals.len;
        } else if (f->frame_type == FT_RAW) size += ch->residuals.len;
    }
    return size;
}

/* ---------------------------------------------------- lossless/encoder.rs -- */
static size_t lpc_order_from_level(uint8_t level) {         /* encoder.rs:289-302 */
    static const uint8_t t[10] = { 0, 2, 4, 4, 6, 8, 8, 10, 12, 12 };
    return level < 10 ? t[level] : 12;
}

static chan_t chan_silence(void) { chan_t c; memset(&c, 0, sizeof c); c.encoding = RE_RICE; return c; } /* types.rs:191-199 */

/* encode_channel_int, encoder.rs:173-217 (+ encode_raw :220-226,
 * try_fixed_predictor :229-251, try_lpc_predictor :254-287).
 * cand (optional) receives the per-candidate report. */
static chan_t encode_channel_int(const int32_t *s, size_t n, size_t max_order, uint8_t level,
                                 size_t *order_used, flo_ref_candidate *cand, int *n_cand) {
    if (n_cand) *n_cand = 0;
    if (n == 0) { *order_used = 0; return chan_silence(); }

    chan_t best = chan_silence(); int have = 0;
    size_t best_size = SIZE_MAX; size_t best_order = 0;
    int32_t *res = (int32_t *)malloc(n * sizeof(int32_t));

    /* Strategy 1: raw PCM, (s as i16).to_le_bytes() */
    {
        chan_t raw = chan_silence(); raw.encoding = RE_RAW;
        bytes_reserve(&raw.residuals, 2 * n);
        for (size_t i = 0; i < n; i++) { uint16_t v = (uint16_t)(int16_t)s[i]; bytes_push(&raw.residuals, (uint8_t)v); bytes_push(&raw.residuals, (uint8_t)(v >> 8)); }
        if (cand) { cand[*n_cand].kind = 0; cand[*n_cand].order = 0; cand[*n_cand].k = 0; cand[*n_cand].size = (int64_t)raw.residuals.len; (*n_cand)++; }
        if (raw.residuals.len < best_size) { best_size = raw.residuals.len; best = raw; have = 1; best_order = 0; }
        else free(raw.residuals.p);
    }
    /* Strategy 2: fixed predictors 0..=min(4, max_order) */
    size_t fmax = max_order < 4 ? max_order : 4;
    for (size_t order = 0; order <= fmax; order++) {
        flo_ref_fixed_predictor_residuals(s, n, (int)order, res);
        uint8_t k = flo_ref_estimate_rice_parameter_i32(res, n);
        bytes_t enc = rice_encode(res, n, k);
        if (cand) { cand[*n_cand].kind = 1; cand[*n_cand].order = (int32_t)order; cand[*n_cand].k = k; cand[*n_cand].size = (int64_t)enc.len; (*n_cand)++; }
        if (enc.len < best_size) {
            if (have) free(best.residuals.p);
            best = chan_silence(); best.shift_bits = (uint8_t)(128 + order); best.encoding = RE_RICE;
            best.rice_parameter = k; best.residuals = enc; have = 1;
            best_size = enc.len; best_order = order;
        } else free(enc.p);
    }
    /* Strategy 3: LPC 5..=max_order */
    if (level >= 3 && max_order > 4) {
        for (size_t order = 5; order <= max_order; order++) {
            if (cand) { cand[*n_cand].kind = 2; cand[*n_cand].order = (int32_t)order; cand[*n_cand].k = 0; cand[*n_cand].size = -1; }
            int ok = 0; bytes_t enc; memset(&enc, 0, sizeof enc);
            int32_t coeffs[12]; uint8_t shift = 0; uint8_t k = 0;
            if (n > order) {
                int64_t ac[13];
                flo_ref_autocorr_int(s, n, (int)order, ac);
                if (flo_ref_levinson_durbin_int(ac, (int)order, coeffs, &shift)) {
                    flo_ref_calc_residuals_int(s, n, coeffs, shift, (int)order, res);
                    int64_t max_res = 0;
                    for (size_t i = 0; i < n; i++) {
                        /* r.abs(): i32::MIN.abs() wraps to i32::MIN in release builds */
                        int32_t a = res[i] == INT32_MIN ? INT32_MIN : (res[i] < 0 ? -res[i] : res[i]);
                        if (i == 0 || a > max_res) max_res = a;
                    }
                    if (max_res <= 1000000) {
                        k = flo_ref_estimate_rice_parameter_i32(res, n);
                        enc = rice_encode(res, n, k);
                        ok = 1;
                    }
                }
            }
            if (cand) { if (ok) { cand[*n_cand].k = k; cand[*n_cand].size = (int64_t)enc.len; } (*n_cand)++; }
            if (ok) {
                if (enc.len < best_size) {
                    if (have) free(best.residuals.p);
                    best = chan_silence(); best.n_coeffs = (uint32_t)order;
                    memcpy(best.coeffs, coeffs, order * sizeof(int32_t));
                    best.shift_bits = shift; best.encoding = RE_RICE; best.rice_parameter = k;
                    best.residuals = enc; have = 1; best_size = enc.len; best_order = order;
                } else free(enc.p);
            }
        }
    }
    free(res);
    *order_used = best_order;
    return best;
}

int flo_ref_channel_candidates(const int32_t *s, size_t n, uint8_t level, flo_ref_candidate *outc) {
    if (level > 9) level = 9;
    size_t used; int nc = 0;
    chan_t c = encode_channel_int(s, n, lpc_order_from_level(level), level, &used, outc, &nc);
    free(c.residuals.p);
    return nc;
}

/* frame assembly from coded-domain channels, encoder.rs:102-127 */
static frame_t encode_frame_from_channels(const int32_t *const *ch, const size_t *ch_len, uint8_t channels,
                                          uint32_t frame_samples, uint8_t flags, uint8_t level) {
    size_t lpc_order = lpc_order_from_level(level);
    frame_t f; memset(&f, 0, sizeof f);
    f.n_channels = channels; f.ch = (chan_t *)calloc(channels ? channels : 1, sizeof(chan_t));
    int all_raw = 1;
    for (uint32_t c = 0; c < channels; c++) {
        size_t used = 0;
        f.ch[c] = encode_channel_int(ch[c], ch_len[c], lpc_order, level, &used, NULL, NULL);
        if (used > 0) all_raw = 0;
    }
    f.frame_type = all_raw ? FT_RAW : ft_from_order(lpc_order);
    f.frame_samples = frame_samples;
    f.flags = flags;
    return f;
}

/* encode_frame, encoder.rs:66-128 */
static frame_t encode_frame(const float *samples, size_t len, uint8_t channels, uint8_t level) {
    size_t C = channels;
    uint32_t num_samples = (uint32_t)(len / C);
    int silent = 1;
    for (size_t i = 0; i < len; i++) if (!(fabsf(samples[i]) < 1e-7f)) { silent = 0; break; }
    if (silent) {
        frame_t f; memset(&f, 0, sizeof f);
        f.frame_type = FT_SILENCE; f.frame_samples = num_samples;
        f.n_channels = channels; f.ch = (chan_t *)calloc(C, sizeof(chan_t));
        for (size_t c = 0; c < C; c++) f.ch[c] = chan_silence();
        return f;
    }
    /* f32 -> i32, deinterleave (skip(ch).step_by(C)): early channels may get one more sample */
    int32_t **cd = (int32_t **)calloc(C, sizeof(int32_t *));
    size_t *cl = (size_t *)calloc(C, sizeof(size_t));
    for (size_t c = 0; c < C; c++) {
        cl[c] = len > c ? (len - c + C - 1) / C : 0;
        cd[c] = (int32_t *)malloc((cl[c] ? cl[c] : 1) * sizeof(int32_t));
        for (size_t i = 0; i < cl[c]; i++) cd[c][i] = flo_ref_f32_to_i32(samples[c + i * C]);
    }
    /* should_use_mid_side, encoder.rs:131-153; to_mid_side :156-170 (zip -> min length) */
    uint8_t flags = 0;
    if (channels == 2) {
        size_t m = cl[0] < cl[1] ? cl[0] : cl[1];
        int64_t var_l = 0, var_r = 0, var_side = 0;
        for (size_t i = 0; i < m; i++) {
            int64_t l = cd[0][i], r = cd[1][i];
            var_l += l * l; var_r += r * r;
            int64_t sd = (int64_t)(int32_t)((uint32_t)cd[0][i] - (uint32_t)cd[1][i]);
            var_side += sd * sd;
        }
        if (var_side < (var_l + var_r) / 2) {
            for (size_t i = 0; i < m; i++) {
                int32_t l = cd[0][i], r = cd[1][i];
                cd[0][i] = (int32_t)((uint32_t)l + (uint32_t)r);
                cd[1][i] = (int32_t)((uint32_t)l - (uint32_t)r);
            }
            cl[0] = cl[1] = m;
            flags |= 0x01;
        }
    }
    frame_t f = encode_frame_from_channels((const int32_t *const *)cd, cl, channels, num_samples, flags, level);
    for (size_t c = 0; c < C; c++) free(cd[c]);
    free(cd); free(cl);
    return f;
}

/* ------------------------------------------------------------- writer.rs -- */
static void write_channel_data(bytes_t *b, const chan_t *ch, uint8_t frame_type) { /* writer.rs:256-301 */
    if (frame_type == FT_SILENCE) return;
    if (frame_type == FT_RAW || frame_type == FT_TRANSFORM) { bytes_extend(b, ch->residuals.p, ch->residuals.len); return; }
    if (ft_is_alpc(frame_type)) {
        bytes_push(b, (uint8_t)ch->n_coeffs);
        for (uint32_t i = 0; i < ch->n_coeffs; i++) bytes_u32(b, (uint32_t)ch->coeffs[i]);
        bytes_push(b, ch->shift_bits);
        bytes_push(b, ch->encoding);
        if (ch->encoding == RE_RICE) bytes_push(b, ch->rice_parameter);
        bytes_extend(b, ch->residuals.p, ch->residuals.len);
    }
}
static void write_frame(bytes_t *b, const frame_t *f) {      /* writer.rs:236-254 */
    bytes_push(b, f->frame_type);
    bytes_u32(b, f->frame_samples);
    bytes_push(b, f->flags);
    for (uint32_t c = 0; c < f->n_channels; c++) {
        bytes_t cb; memset(&cb, 0, sizeof cb);
        write_channel_data(&cb, &f->ch[c], f->frame_type);
        bytes_u32(b, (uint32_t)cb.len);
        bytes_extend(b, cb.p, cb.len);
        free(cb.p);
    }
}

/* Writer::write / write_ex / write_header_ex / build_toc_chunk, writer.rs:16-224 */
static bytes_t write_file(uint32_t sample_rate, uint8_t channels, uint8_t bit_depth, uint8_t level,
                          const frame_t *frames, size_t n_frames, const uint8_t *meta, size_t meta_len) {
    bytes_t data; memset(&data, 0, sizeof data);
    for (size_t i = 0; i < n_frames; i++) write_frame(&data, &frames[i]);
    uint64_t toc_size = 4 + (uint64_t)n_frames * 20;
    uint32_t crc = flo_ref_crc32(data.p, data.len);

    bytes_t toc; memset(&toc, 0, sizeof toc);
    bytes_u32(&toc, (uint32_t)n_frames);
    uint64_t byte_offset = 0, cum = 0, total_samples = 0;
    for (size_t i = 0; i < n_frames; i++) {
        uint32_t fs = (uint32_t)frame_byte_size(&frames[i]);
        bytes_u32(&toc, (uint32_t)i);
        bytes_u64(&toc, byte_offset);
        bytes_u32(&toc, fs);
        bytes_u32(&toc, (uint32_t)(cum * 1000 / (uint64_t)sample_rate));
        byte_offset += fs;
        cum += frames[i].frame_samples;
        total_samples += frames[i].frame_samples;
    }

    bytes_t out; memset(&out, 0, sizeof out);
    static const uint8_t magic[4] = { 0x46, 0x4c, 0x4f, 0x21 };
    bytes_extend(&out, magic, 4);
    bytes_push(&out, 1); bytes_push(&out, 2);               /* version 1.2, types.rs:12-13 */
    bytes_u16(&out, 0);                                      /* flags (lossless) */
    bytes_u32(&out, sample_rate);
    bytes_push(&out, channels);
    bytes_push(&out, bit_depth);
    bytes_u64(&out, total_samples);
    bytes_push(&out, level);
    bytes_push(&out, 0); bytes_push(&out, 0); bytes_push(&out, 0);
    bytes_u32(&out, crc);
    bytes_u64(&out, 66);                                     /* HEADER_SIZE, types.rs:9 */
    bytes_u64(&out, toc_size);
    bytes_u64(&out, (uint64_t)data.len);
    bytes_u64(&out, 0);
    bytes_u64(&out, (uint64_t)meta_len);
    bytes_extend(&out, toc.p, toc.len);
    bytes_extend(&out, data.p, data.len);
    bytes_extend(&out, meta, meta_len);
    free(toc.p); free(data.p);
    return out;
}

/* Encoder::encode + encode_frames, encoder.rs:32-64 */
int flo_ref_encode(const float *samples, size_t n, uint32_t sample_rate, uint8_t channels, uint8_t bit_depth,
                   uint8_t level, const uint8_t *meta, size_t meta_len, uint8_t **out, size_t *out_len) {
    if (channels == 0) { set_err("channels == 0 (reference panics: division by zero)"); return -1; }
    if (sample_rate == 0) { set_err("sample_rate == 0 (reference panics: division by zero)"); return -1; }
    if (level > 9) level = 9;                                /* with_compression: level.min(9) */
    size_t C = channels, spf = sample_rate;
    size_t total = n / C;
    size_t n_frames = (total + spf - 1) / spf;
    frame_t *frames = (frame_t *)calloc(n_frames ? n_frames : 1, sizeof(frame_t));
    for (size_t i = 0; i < n_frames; i++) {
        size_t start = i * spf * C;
        size_t end = (i + 1) * spf * C; if (end > n) end = n;
        frames[i] = encode_frame(samples + start, end - start, channels, level);
    }
    bytes_t b = write_file(sample_rate, channels, bit_depth, level, frames, n_frames, meta, meta_len);
    for (size_t i = 0; i < n_frames; i++) frame_free(&frames[i]);
    free(frames);
    *out = b.p; *out_len = b.len;
    return 0;
}

int flo_ref_encode_pcm16(const int16_t *pcm, size_t n, uint32_t sample_rate, uint8_t channels, uint8_t bit_depth,
                         uint8_t level, const uint8_t *meta, size_t meta_len, uint8_t **out, size_t *out_len) {
    float *f = (float *)malloc((n ? n : 1) * sizeof(float));
    const float scale = 1.0f / 32768.0f;                     /* reflo/src/audio.rs:248 */
    for (size_t i = 0; i < n; i++) f[i] = (float)pcm[i] * scale;
    int rc = flo_ref_encode(f, n, sample_rate, channels, bit_depth, level, meta, meta_len, out, out_len);
    free(f);
    return rc;
}

int flo_ref_encode_frame_i32(const int32_t *const *ch, const size_t *ch_len, uint8_t channels, uint32_t frame_samples,
                             uint8_t flags, uint8_t level, uint8_t **out, size_t *out_len) {
    if (level > 9) level = 9;
    frame_t f = encode_frame_from_channels(ch, ch_len, channels, frame_samples, flags, level);
    bytes_t b; memset(&b, 0, sizeof b);
    write_frame(&b, &f);
    if (b.len != frame_byte_size(&f)) { set_err("frame_byte_size mismatch"); frame_free(&f); free(b.p); return -1; }
    frame_free(&f);
    *out = b.p; *out_len = b.len;
    return 0;
}

/* ------------------------------------------------------------- reader.rs -- */
struct flo_ref_file {
    uint8_t version_major, version_minor; uint16_t flags;
    uint32_t sample_rate; uint8_t channels, bit_depth; uint64_t total_samples;
    uint8_t level; uint32_t crc; uint64_t header_size, toc_size, data_size, extra_size, meta_size;
    uint64_t data_offset;
    uint32_t n_toc; uint64_t *toc_off; uint32_t *toc_size_e, *toc_ts;
    uint32_t n_frames; frame_t *frames;
};

typedef struct { const uint8_t *d; size_t len, pos; int fail; } cur_t;
static uint8_t c_u8(cur_t *c) { if (c->pos >= c->len) { c->fail = 1; return 0; } return c->d[c->pos++]; }
static uint64_t c_le(cur_t *c, int n) {
    if (c->pos + (size_t)n > c->len) { c->fail = 1; return 0; }
    uint64_t v = 0; for (int i = 0; i < n; i++) v |= (uint64_t)c->d[c->pos + i] << (8 * i);
    c->pos += (size_t)n; return v;
}
static void c_skip(cur_t *c, size_t n) { c->pos = c->pos + n < c->len ? c->pos + n : c->len; }
static int c_bytes(cur_t *c, size_t n, bytes_t *out) {
    if (c->pos + n > c->len) { c->fail = 1; return 0; }
    bytes_extend(out, c->d + c->pos, n); c->pos += n; return 1;
}

/* read_channel_data, reader.rs:168-247 */
static int read_channel_data(cur_t *c, uint8_t ft, size_t frame_samples, size_t ch_end, chan_t *out) {
    *out = chan_silence();
    if (frame_samples > 2000000) { set_err("Invalid frame: too many samples"); return 0; }
    if (ft == FT_SILENCE) return 1;
    if (ft == FT_RAW) {
        size_t need = frame_samples * 2, avail = ch_end > c->pos ? ch_end - c->pos : 0;
        out->encoding = RE_RAW;
        return c_bytes(c, need < avail ? need : avail, &out->residuals);
    }
    if (ft == FT_TRANSFORM) {
        size_t rem = ch_end > c->pos ? ch_end - c->pos : 0;
        out->encoding = RE_RAW;
        return rem ? c_bytes(c, rem, &out->residuals) : 1;
    }
    if (ft_is_alpc(ft)) {
        size_t order = c_u8(c);
        if (c->fail) return 0;
        if (order > 12) { set_err("Invalid LPC order"); return 0; }
        for (size_t i = 0; i < order; i++) {
            if (c->pos + 4 > ch_end) break;
            out->coeffs[out->n_coeffs++] = (int32_t)(uint32_t)c_le(c, 4);
        }
        out->shift_bits = c_u8(c);
        uint8_t eb = c_u8(c);
        out->encoding = eb == 0 ? RE_RICE : (eb == 1 ? RE_GOLOMB : RE_RAW);
        out->rice_parameter = out->encoding == RE_RICE ? c_u8(c) : 0;
        if (c->fail) return 0;
        size_t rem = ch_end > c->pos ? ch_end - c->pos : 0;
        return rem ? c_bytes(c, rem, &out->residuals) : 1;
    }
    return 1;
}

flo_ref_file *flo_ref_parse(const uint8_t *data, size_t len) { /* reader.rs:16-166 */
    cur_t c = { data, len, 0, 0 };
    if (len < 4 || memcmp(data, "FLO!", 4) != 0) { set_err("Invalid flo file: bad magic"); return NULL; }
    c.pos = 4;
    flo_ref_file *f = (flo_ref_file *)calloc(1, sizeof *f);
    f->version_major = c_u8(&c); f->version_minor = c_u8(&c); f->flags = (uint16_t)c_le(&c, 2);
    f->sample_rate = (uint32_t)c_le(&c, 4); f->channels = c_u8(&c); f->bit_depth = c_u8(&c);
    f->total_samples = c_le(&c, 8); f->level = c_u8(&c); c_skip(&c, 3);
    f->crc = (uint32_t)c_le(&c, 4); f->header_size = c_le(&c, 8); f->toc_size = c_le(&c, 8);
    f->data_size = c_le(&c, 8); f->extra_size = c_le(&c, 8); f->meta_size = c_le(&c, 8);
    if (c.fail) { set_err("Unexpected end of file"); flo_ref_file_free(f); return NULL; }
    if (f->toc_size >= 4) {
        uint32_t ne = (uint32_t)c_le(&c, 4);
        if (ne > 100000) { set_err("Invalid TOC: too many entries"); flo_ref_file_free(f); return NULL; }
        f->n_toc = ne;
        f->toc_off = (uint64_t *)calloc(ne ? ne : 1, sizeof(uint64_t));
        f->toc_size_e = (uint32_t *)calloc(ne ? ne : 1, sizeof(uint32_t));
        f->toc_ts = (uint32_t *)calloc(ne ? ne : 1, sizeof(uint32_t));
        for (uint32_t i = 0; i < ne; i++) {
            (void)c_le(&c, 4);
            f->toc_off[i] = c_le(&c, 8); f->toc_size_e[i] = (uint32_t)c_le(&c, 4); f->toc_ts[i] = (uint32_t)c_le(&c, 4);
        }
        if (c.fail) { set_err("Unexpected end of file"); flo_ref_file_free(f); return NULL; }
    }
    size_t data_start = c.pos, data_end = c.pos + (size_t)f->data_size;
    f->data_offset = data_start;
    f->frames = (frame_t *)calloc(f->n_toc ? f->n_toc : 1, sizeof(frame_t));
    for (uint32_t i = 0; i < f->n_toc; i++) {
        size_t fstart = data_start + (size_t)f->toc_off[i];
        if (fstart >= data_end) break;
        c.pos = fstart;
        size_t fend = fstart + f->toc_size_e[i];
        frame_t *fr = &f->frames[f->n_frames];
        fr->frame_type = c_u8(&c); fr->frame_samples = (uint32_t)c_le(&c, 4); fr->flags = c_u8(&c);
        if (c.fail) { set_err("Unexpected end of file"); flo_ref_file_free(f); return NULL; }
        uint32_t nch = fr->frame_type == FT_TRANSFORM ? 1 : f->channels;
        fr->ch = (chan_t *)calloc(nch ? nch : 1, sizeof(chan_t));
        f->n_frames++;
        for (uint32_t k = 0; k < nch; k++) {
            size_t ch_size = (size_t)c_le(&c, 4);
            size_t ch_end = c.pos + ch_size;
            if (c.fail || !read_channel_data(&c, fr->frame_type, fr->frame_samples, ch_end, &fr->ch[k])) {
                if (c.fail) set_err("Unexpected end of file");
                fr->n_channels = k + 1;
                flo_ref_file_free(f); return NULL;
            }
            fr->n_channels = k + 1;
            c.pos = ch_end;
        }
        c.pos = fend;
    }
    /* reader.rs:35-43: cursor at the end of DATA, EXTRA skipped (clamped to the file), META read in full */
    c.pos = data_end;
    c_skip(&c, (size_t)f->extra_size);
    if (c.pos > len) c.pos = len;
    if ((size_t)f->meta_size > len - c.pos) { set_err("Unexpected end of file"); flo_ref_file_free(f); return NULL; }
    return f;
}

void flo_ref_file_free(flo_ref_file *f) {
    if (!f) return;
    for (uint32_t i = 0; i < f->n_frames; i++) frame_free(&f->frames[i]);
    free(f->frames); free(f->toc_off); free(f->toc_size_e); free(f->toc_ts); free(f);
}
uint32_t flo_ref_file_sample_rate(const flo_ref_file *f) { return f->sample_rate; }
uint8_t  flo_ref_file_channels(const flo_ref_file *f) { return f->channels; }
uint8_t  flo_ref_file_bit_depth(const flo_ref_file *f) { return f->bit_depth; }
uint8_t  flo_ref_file_level(const flo_ref_file *f) { return f->level; }
uint64_t flo_ref_file_total_samples(const flo_ref_file *f) { return f->total_samples; }
uint32_t flo_ref_file_crc32(const flo_ref_file *f) { return f->crc; }
uint64_t flo_ref_file_data_offset(const flo_ref_file *f) { return f->data_offset; }
uint64_t flo_ref_file_data_size(const flo_ref_file *f) { return f->data_size; }
uint64_t flo_ref_file_meta_size(const flo_ref_file *f) { return f->meta_size; }
uint32_t flo_ref_file_num_frames(const flo_ref_file *f) { return f->n_frames; }
void flo_ref_file_frame_info(const flo_ref_file *f, uint32_t i, uint8_t *type, uint32_t *samples, uint8_t *flags,
                             uint64_t *byte_offset, uint32_t *frame_size, uint32_t *timestamp_ms) {
    const frame_t *fr = &f->frames[i];
    if (type) *type = fr->frame_type;
    if (samples) *samples = fr->frame_samples;
    if (flags) *flags = fr->flags;
    if (byte_offset) *byte_offset = f->toc_off[i];
    if (frame_size) *frame_size = f->toc_size_e[i];
    if (timestamp_ms) *timestamp_ms = f->toc_ts[i];
}
void flo_ref_file_channel_info(const flo_ref_file *f, uint32_t i, uint32_t c, uint32_t *n_coeffs, uint8_t *shift_bits,
                               uint8_t *encoding, uint8_t *k, uint64_t *residual_bytes, int32_t *coeffs12) {
    const chan_t *ch = &f->frames[i].ch[c];
    if (n_coeffs) *n_coeffs = ch->n_coeffs;
    if (shift_bits) *shift_bits = ch->shift_bits;
    if (encoding) *encoding = ch->encoding;
    if (k) *k = ch->rice_parameter;
    if (residual_bytes) *residual_bytes = ch->residuals.len;
    if (coeffs12) memcpy(coeffs12, ch->coeffs, sizeof ch->coeffs);
}

/* --------------------------------------------------- lossless/decoder.rs -- */
static void reconstruct_fixed(int order, const int32_t *r, size_t rlen, size_t target, int32_t *s) { /* decoder.rs:187-270 */
    size_t m = rlen < target ? rlen : target;
    for (size_t i = 0; i < target; i++) s[i] = 0;
    if (rlen == 0) return;
    if (order < 1 || order > 4) { memcpy(s, r, m * sizeof(int32_t)); return; }
    for (size_t i = 0; i < m; i++) {
        /* for i < order the order-i predictor is used (decoder.rs:208-263) */
        int o = (size_t)order < i ? order : (int)i;
        int64_t pred = 0;
        switch (o) {
        case 1: pred = s[i - 1]; break;
        case 2: pred = 2 * (int64_t)s[i - 1] - (int64_t)s[i - 2]; break;
        case 3: pred = 3 * (int64_t)s[i - 1] - 3 * (int64_t)s[i - 2] + (int64_t)s[i - 3]; break;
        case 4: pred = 4 * (int64_t)s[i - 1] - 6 * (int64_t)s[i - 2] + 4 * (int64_t)s[i - 3] - (int64_t)s[i - 4]; break;
        default: pred = 0; break;
        }
        s[i] = (int32_t)((uint32_t)r[i] + (uint32_t)(int32_t)pred);
    }
}

static void reconstruct_lpc_int(const int32_t *coeffs, const int32_t *r, uint8_t shift, size_t order, size_t target, int32_t *s) { /* decoder.rs:152-184 */
    for (size_t i = 0; i < target; i++) s[i] = 0;
    size_t warm = order < target ? order : target;
    for (size_t i = 0; i < warm; i++) s[i] = r[i];
    for (size_t i = order; i < target; i++) {
        int64_t pred = 0;
        for (size_t j = 0; j < order; j++) pred += (int64_t)coeffs[j] * (int64_t)s[i - j - 1];
        s[i] = (int32_t)((uint32_t)(int32_t)(pred >> shift) + (uint32_t)r[i]);
    }
}

/* decode_channel_int, decoder.rs:92-149 */
static void decode_channel_int(const chan_t *ch, size_t frame_samples, int32_t *out) {
    int has_coeffs = ch->n_coeffs > 0, has_res = ch->residuals.len > 0;
    if (!has_coeffs && has_res && ch->shift_bits >= 128) {
        int32_t *r = (int32_t *)malloc((frame_samples ? frame_samples : 1) * sizeof(int32_t));
        flo_ref_rice_decode_i32(ch->residuals.p, ch->residuals.len, ch->rice_parameter, frame_samples, r);
        reconstruct_fixed(ch->shift_bits - 128, r, frame_samples, frame_samples, out);
        free(r); return;
    }
    if (has_coeffs) {
        int32_t *r = (int32_t *)malloc((frame_samples ? frame_samples : 1) * sizeof(int32_t));
        flo_ref_rice_decode_i32(ch->residuals.p, ch->residuals.len, ch->rice_parameter, frame_samples, r);
        reconstruct_lpc_int(ch->coeffs, r, ch->shift_bits, ch->n_coeffs, frame_samples, out);
        free(r); return;
    }
    if (has_res) {
        size_t k = 0;
        for (size_t i = 0; i + 1 < ch->residuals.len && k < frame_samples; i += 2)
            out[k++] = (int32_t)(int16_t)((uint16_t)ch->residuals.p[i] | ((uint16_t)ch->residuals.p[i + 1] << 8));
        /* note: the reference pushes every complete 2-byte chunk; the reader never hands it more than 2*frame_samples */
        while (k < frame_samples) out[k++] = 0;
        return;
    }
    for (size_t i = 0; i < frame_samples; i++) out[i] = 0;
}

int flo_ref_file_decode_frame_coded(const flo_ref_file *f, uint32_t i, int32_t *out) {
    const frame_t *fr = &f->frames[i];
    for (uint32_t c = 0; c < fr->n_channels; c++) decode_channel_int(&fr->ch[c], fr->frame_samples, out + (size_t)c * fr->frame_samples);
    return 0;
}

/* decode_file, decoder.rs:21-73 (+ decode_mid_side :75-89) */
int flo_ref_decode_i32(const uint8_t *data, size_t len, int32_t **out, size_t *out_n) {
    flo_ref_file *f = flo_ref_parse(data, len);
    if (!f) return -1;
    size_t C = f->channels;
    size_t total = 0;
    for (uint32_t i = 0; i < f->n_frames; i++) total += f->frames[i].frame_samples;
    int32_t *pl = (int32_t *)calloc((C * total) ? C * total : 1, sizeof(int32_t));   /* planar [C][total] */
    size_t *fill = (size_t *)calloc(C ? C : 1, sizeof(size_t));
    for (uint32_t i = 0; i < f->n_frames; i++) {
        const frame_t *fr = &f->frames[i];
        size_t n = fr->frame_samples;
        int32_t *tmp = (int32_t *)malloc((fr->n_channels * n ? fr->n_channels * n : 1) * sizeof(int32_t));
        for (uint32_t c = 0; c < fr->n_channels; c++) decode_channel_int(&fr->ch[c], n, tmp + (size_t)c * n);
        int ms = C == 2 && (fr->flags & 1) && fr->n_channels == 2;
        if (ms) {
            for (size_t j = 0; j < n; j++) {
                int32_t m = tmp[j], s = tmp[n + j];
                int32_t a = (int32_t)((uint32_t)m + (uint32_t)s), b = (int32_t)((uint32_t)m - (uint32_t)s);
                pl[0 * total + fill[0] + j] = a / 2;          /* Rust `/`: truncates toward zero */
                pl[1 * total + fill[1] + j] = b / 2;
            }
            fill[0] += n; fill[1] += n;
        } else {
            for (uint32_t c = 0; c < fr->n_channels && c < C; c++) {
                memcpy(pl + c * total + fill[c], tmp + (size_t)c * n, n * sizeof(int32_t));
                fill[c] += n;
            }
        }
        free(tmp);
    }
    size_t max_len = 0;
    for (size_t c = 0; c < C; c++) if (fill[c] > max_len) max_len = fill[c];
    int32_t *il = (int32_t *)calloc((max_len * C) ? max_len * C : 1, sizeof(int32_t));
    for (size_t i = 0; i < max_len; i++)
        for (size_t c = 0; c < C; c++) il[i * C + c] = i < fill[c] ? pl[c * total + i] : 0;
    free(pl); free(fill);
    *out = il; *out_n = max_len * C;
    flo_ref_file_free(f);
    return 0;
}

int flo_ref_decode(const uint8_t *data, size_t len, float **out, size_t *out_n) {
    int32_t *il; size_t n;
    if (flo_ref_decode_i32(data, len, &il, &n) != 0) return -1;
    float *o = (float *)malloc((n ? n : 1) * sizeof(float));
    for (size_t i = 0; i < n; i++) o[i] = flo_ref_i32_to_f32(il[i]);
    free(il);
    *out = o; *out_n = n;
    return 0;
}
