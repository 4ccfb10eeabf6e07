/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's EBU R128 integrated loudness
 * (libflo/src/core/ebu_r128.rs), the value libflo::encode() stores as loudness_profile[0].lufs
 * (libflo/src/lib.rs:256-268).  Sequential f64, every product and sum rounded on its own like the
 * reference's plain arithmetic (the library is built with -ffp-contract=off), filter states carried
 * over the whole channel.  Only tests/ may call this.
 *
 * PARITY UNPINNED: the reference's tests (libflo/tests/rust/loudness_tests.rs) assert ranges only
 * (e.g. a 440 Hz sine of amplitude 0.5 lies between -25 and -5 LUFS); tests/test_oracle_golden.py
 * replays them and adds the BS.1770 calibration point (a full-scale 997 Hz sine in one channel
 * reads -3.01 LKFS).  LRA, true peak and sample peak of LoudnessMetrics are not restated: encode()
 * does not store them. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

typedef struct { double b0, b1, b2, a1, a2, z1, z2; } biquad;          /* ebu_r128.rs:19-27 */

static double biquad_process(biquad *f, double x) {                      /* ebu_r128.rs:43-48 */
    const double y = f->b0 * x + f->z1;
    f->z1 = f->b1 * x - f->a1 * y + f->z2;
    f->z2 = f->b2 * x - f->a2 * y;
    return y;
}

/* KWeighting::new, ebu_r128.rs:58-102; out: shelf b0 b1 b2 a1 a2, then high-pass b0 b1 b2 a1 a2 */
void flo_ref_kweighting_coeffs(double sample_rate, double out[10]) {
    const double pi = 3.14159265358979323846264338327950288;
    const double f0 = 1681.974450955533, g_db = 3.999843853973347, q = 0.7071752369554196;
    const double k = tan(pi * f0 / sample_rate);
    const double vh = pow(10.0, g_db / 20.0);
    const double vb = pow(vh, 0.4996667741545416);
    const double a0 = 1.0 + k / q + k * k;
    out[0] = (vh + vb * k / q + k * k) / a0;
    out[1] = 2.0 * (k * k - vh) / a0;
    out[2] = (vh - vb * k / q + k * k) / a0;
    out[3] = 2.0 * (k * k - 1.0) / a0;
    out[4] = (1.0 - k / q + k * k) / a0;
    const double f0_hp = 38.13547087602444, q_hp = 0.5003270373238773;
    const double k_hp = tan(pi * f0_hp / sample_rate);
    const double a0_hp = 1.0 + k_hp / q_hp + k_hp * k_hp;
    out[5] = 1.0; out[6] = -2.0; out[7] = 1.0;
    out[8] = 2.0 * (k_hp * k_hp - 1.0) / a0_hp;
    out[9] = (1.0 - k_hp / q_hp + k_hp * k_hp) / a0_hp;
}

/* compute_ebu_r128_loudness(...).integrated_lufs, ebu_r128.rs:182-313.
 * block_energies_out (optional, malloc'd, *n_blocks entries) lets the tests look at the 400 ms blocks. */
double flo_ref_r128_integrated(const float *samples, size_t n_interleaved, uint8_t channels, uint32_t sample_rate,
                               double **block_energies_out, size_t *n_blocks) {
    if (block_energies_out) *block_energies_out = NULL;
    if (n_blocks) *n_blocks = 0;
    if (n_interleaved == 0 || channels == 0) return -23.0;               /* :187-194 */
    const double sr = (double)sample_rate;
    const size_t hop = (size_t)round(sr * 0.1);                          /* :197 (f64::round: half away from zero) */
    const size_t block = hop * 4;                                        /* :198 */
    const size_t frames = n_interleaved / channels;                      /* :201 */
    double co[10];
    flo_ref_kweighting_coeffs(sr, co);
    double *kw = (double *)malloc(sizeof(double) * (frames ? frames : 1) * channels);
    for (size_t ch = 0; ch < channels; ch++) {                           /* :221-230 */
        biquad shelf = {co[0], co[1], co[2], co[3], co[4], 0.0, 0.0};
        biquad hp = {co[5], co[6], co[7], co[8], co[9], 0.0, 0.0};
        for (size_t i = 0; i < frames; i++)
            kw[ch * frames + i] = biquad_process(&hp, biquad_process(&shelf, (double)samples[i * channels + ch]));
    }
    size_t cap = frames / (hop ? hop : 1) + 8, nb = 0;
    double *be = (double *)malloc(sizeof(double) * cap);
    size_t start = 0;
    while (start < frames) {                                             /* :236-265 */
        size_t end = start + block < frames ? start + block : frames;
        if (end <= start) break;
        double energy = 0.0;
        const size_t len = end - start;
        for (size_t ch = 0; ch < channels; ch++) {
            double sum_sq = 0.0;
            for (size_t i = start; i < end; i++) sum_sq += kw[ch * frames + i] * kw[ch * frames + i];
            energy += sum_sq / (double)len;
        }
        if (nb == cap) { cap *= 2; be = (double *)realloc(be, sizeof(double) * cap); }
        be[nb++] = energy;
        if (end == frames) break;
        start += hop;
        if (hop == 0) break;                                             /* sample_rate < 5: the reference loops forever */
    }
    free(kw);
    double result = -23.0;
    if (nb) {
        const double abs_gate = pow(10.0, (-70.0 + 0.691) / 10.0);       /* :278-279 */
        double sum_e = 0.0; size_t cnt = 0;
        for (size_t i = 0; i < nb; i++) if (be[i] >= abs_gate) { sum_e += be[i]; cnt++; }
        if (cnt) {                                                       /* :297-313 */
            const double ungated = -0.691 + 10.0 * log10(sum_e / (double)cnt);
            const double rel_gate = pow(10.0, (ungated - 10.0 + 0.691) / 10.0);
            double s2 = 0.0; size_t c2 = 0;
            for (size_t i = 0; i < nb; i++) if (be[i] >= abs_gate && be[i] >= rel_gate) { s2 += be[i]; c2++; }
            result = c2 ? -0.691 + 10.0 * log10(s2 / (double)c2) : ungated;
        }
    }
    if (block_energies_out) { *block_energies_out = be; if (n_blocks) *n_blocks = nb; }
    else free(be);
    return result;
}
