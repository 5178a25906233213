/* tf_oracle.c - CPU restatement of the temporal filter's planewise weighting (TEST INFRASTRUCTURE: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu arm may use it).
 *
 * Follows Source/Lib/Encoder/Codec/EbTemporalFiltering.c:
 *   svt_av1_apply_temporal_filter_planewise_c      :643-811   (8 bit)
 *   svt_av1_apply_temporal_filter_planewise_hbd_c  :829-1017  (10 bit: sums >> 2*(bd-8), block errors >> 4)
 *   calculate_squared_errors[_highbd]              :525-556
 *   apply_filtering_central[_highbd]               :551-621   (the centre frame: weight 1000 everywhere)
 *   get_final_filtered_pixels                      :1943-2050 (round(accum / count) and the filtered SSE)
 * The reference mixes integer window sums with a short double / float chain per sample:
 *     window_error = (double)sum / n;  combined = (5 * window_error + block_error) / 6;
 *     scaled = min(combined * d_factor / (2 n_decay^2) / 1 / 1, 7);  weight = (int)(expf((float)-scaled) * 1000)
 * IEEE double +, *, / are reproducible anywhere; expf is not specified bit for bit, so it is restated here as the
 * algorithm glibc (2.27 and later: sysdeps/ieee754/flt-32/e_expf.c, the 32-entry table method of Szabolcs Nagy) uses,
 * and tests/test_oracle_tf.py compares orc_expf with the host's libm for EVERY float in [-8, -0] (1.09e9 values; the
 * filter only evaluates [-7, -0]).  Parity of the GPU path rests on that exhaustive identity, not on a tolerance. */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* T[i] = bits(2^(i/32)) - (i << 47), 2^(i/32) correctly rounded (generated with 200-bit arithmetic, tools/gen_exp2f_table.py) */
static const uint64_t k_exp2f_tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull,
    0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull,
    0x3feedea64c123422ull, 0x3feece086061892dull, 0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull,
    0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, 0x3feee89f995ad3adull,
    0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
    0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

const uint64_t *orc_exp2f_table(void) { return k_exp2f_tab; }

/* expf for finite x in the range where no overflow / underflow handling is needed (|x| < 87); every step is one
 * rounded double operation, written so that the compiler cannot fuse or reorder them (the GPU kernel uses the same
 * sequence with explicit round-to-nearest intrinsics). */
float orc_expf(float x) {
    const double inv_ln2_n = 0x1.71547652b82fep+0 * 32, shift = 0x1.8p52;
    const double c0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, c1 = 0x1.ebfce50fac4f3p-3 / 32 / 32, c2 = 0x1.62e42ff0c52d6p-1 / 32;
    volatile double z = inv_ln2_n * (double)x;
    volatile double kd = z + shift;
    uint64_t ki;
    double kdv = kd;
    memcpy(&ki, &kdv, 8);
    volatile double kd2 = kd - shift;
    volatile double r = z - kd2;
    uint64_t t = k_exp2f_tab[ki % 32] + (ki << 47);
    double s;
    memcpy(&s, &t, 8);
    volatile double a = c0 * r;
    volatile double zz = a + c1;
    volatile double r2 = r * r;
    volatile double b = c2 * r;
    volatile double y = b + 1.0;
    volatile double c = zz * r2;
    volatile double y2 = c + y;
    volatile double y3 = y2 * s;
    return (float)y3;
}

/* count of floats in [lo_bits, hi_bits] (as sign-magnitude bit patterns of negative floats) whose orc_expf differs
 * from the host libm's expf; the second counter is the same for the filter weight (int)(e * 1000). */
long orc_expf_mismatches(uint32_t lo_bits, uint32_t hi_bits, long *weight_mismatches) {
    long bad = 0, badw = 0;
#pragma omp parallel for reduction(+ : bad, badw) schedule(static)
    for (uint64_t u = lo_bits; u <= hi_bits; u++) {
        uint32_t v = (uint32_t)u;
        float x;
        memcpy(&x, &v, 4);
        const float e = expf(x), m = orc_expf(x);
        if (memcmp(&e, &m, 4)) bad++;
        if ((int)(e * 1000) != (int)(m * 1000)) badw++;
    }
    if (weight_mismatches) *weight_mismatches = badw;
    return bad;
}

/* xor-rotate checksum of orc_expf's bit patterns over a range: the GPU test compares its kernel's checksum of the same
 * range (all 1.09e9 floats of [-8, -0] without moving 4 GB). */
uint64_t orc_expf_checksum(uint32_t lo_bits, uint32_t hi_bits) {
    uint64_t acc = 0;
#pragma omp parallel for reduction(+ : acc) schedule(static)
    for (uint64_t u = lo_bits; u <= hi_bits; u++) {
        uint32_t v = (uint32_t)u, o;
        float x;
        memcpy(&x, &v, 4);
        const float m = orc_expf(x);
        memcpy(&o, &m, 4);
        acc += (uint64_t)o * (uint64_t)(2 * (v & 0xffff) + 1);
    }
    return acc;
}

/* The per-plane denominator 2 * n_decay^2 with n_decay = decay_control * (0.7 + log1p(noise)) (:712, :785, :800). */
void orc_tf_den(int decay_control, const double *noise_levels, double den[3]) {
    for (int p = 0; p < 3; p++) {
        const double n_decay = (double)decay_control * (0.7 + log1p(noise_levels[p]));
        den[p] = 2 * n_decay * n_decay;
    }
}

/* block_error and d_factor of the four 16x16 quadrants of one 32x32 block (:696-734, :884-924) from the ME results the
 * reference keeps in MeContext (tf_32x32_block_split_flag, tf_16x16/32x32_block_error, tf_16x16/32x32_mv_x/y). */
void orc_tf_block_factors(int split, const uint64_t err16[4], uint64_t err32, const int16_t mvx16[4], const int16_t mvy16[4],
                          int16_t mvx32, int16_t mvy32, int min_frame_size, int hbd, double block_error[4], double d_factor[4]) {
    for (int q = 0; q < 4; q++) {
        if (split)
            block_error[q] = (double)(hbd ? err16[q] >> 4 : err16[q]) / 256;
        else
            block_error[q] = (double)(hbd ? err32 >> 4 : err32) / 1024;
        const int16_t col = split ? mvx16[q] : mvx32, row = split ? mvy16[q] : mvy32;
        const float distance = sqrtf(powf(row, 2) + powf(col, 2));
        const double thr = min_frame_size * 0.1 > 1 ? min_frame_size * 0.1 : 1;
        const double dt = (double)thr;
        d_factor[q] = distance / dt > 1 ? distance / dt : 1;
    }
}

int orc_tf_weight(uint64_t sum, int n, double block_error, double d_factor, double den) {
    volatile double window_error = (double)sum / n;
    volatile double w5 = 5 * window_error;
    volatile double num = w5 + block_error;
    volatile double combined = num / 6;
    volatile double a = combined * d_factor;
    volatile double b = a / den;
    double scaled = b < 7 ? b : 7;
    return (int)(orc_expf((float)(-scaled)) * 1000);
}

static inline int px(const void *p, int idx, int hbd) { return hbd ? ((const uint16_t *)p)[idx] : ((const uint8_t *)p)[idx]; }

/* One block (the reference calls it per 32x32 luma block; any even size up to 64x64 works).  accum / count use the
 * PREDICTION's strides, as in the reference (k = i * y_pre_stride + j). */
void orc_tf_planewise(const void *y_src, int y_src_stride, const void *y_pre, int y_pre_stride, const void *u_src, const void *v_src,
                      int uv_src_stride, const void *u_pre, const void *v_pre, int uv_pre_stride, unsigned bw, unsigned bh, int ss_x,
                      int ss_y, const double den[3], const double block_error[4], const double d_factor[4], int chroma, int bit_depth,
                      uint32_t *y_accum, uint16_t *y_count, uint32_t *u_accum, uint16_t *u_count, uint32_t *v_accum, uint16_t *v_count) {
    const int hbd = bit_depth > 8, sh = (bit_depth - 8) * 2;
    const unsigned uw = bw >> ss_x, uh = bh >> ss_y;
    uint32_t *yd = calloc(4096 * 3, sizeof(uint32_t)), *ud = yd + 4096, *vd = yd + 8192;
    for (unsigned i = 0; i < bh; i++)
        for (unsigned j = 0; j < bw; j++) {
            const int d = px(y_src, i * y_src_stride + j, hbd) - px(y_pre, i * y_pre_stride + j, hbd);
            yd[i * bw + j] = (uint32_t)(d * d);
        }
    if (chroma)
        for (unsigned i = 0; i < uh; i++)
            for (unsigned j = 0; j < uw; j++) {
                const int du = px(u_src, i * uv_src_stride + j, hbd) - px(u_pre, i * uv_pre_stride + j, hbd);
                const int dv = px(v_src, i * uv_src_stride + j, hbd) - px(v_pre, i * uv_pre_stride + j, hbd);
                ud[i * uw + j] = (uint32_t)(du * du);
                vd[i * uw + j] = (uint32_t)(dv * dv);
            }
    for (unsigned i = 0; i < bh; i++)
        for (unsigned j = 0; j < bw; j++) {
            uint64_t sum = 0;
            for (int dy = -2; dy <= 2; dy++)
                for (int dx = -2; dx <= 2; dx++) {
                    int r = (int)i + dy, c = (int)j + dx;
                    r = r < 0 ? 0 : r > (int)bh - 1 ? (int)bh - 1 : r;
                    c = c < 0 ? 0 : c > (int)bw - 1 ? (int)bw - 1 : c;
                    sum += yd[r * (int)bw + c];
                }
            const int q = (i >= bh / 2) * 2 + (j >= bw / 2);
            int w = orc_tf_weight(sum >> sh, 25, block_error[q], d_factor[q], den[0]);
            unsigned k = i * y_pre_stride + j;
            y_count[k] += (uint16_t)w;
            y_accum[k] += (uint32_t)(w * px(y_pre, k, hbd));
            if (chroma && !(i & ss_y) && !(j & ss_x)) {
                const int ur = i >> ss_y, uc = j >> ss_x;
                uint64_t s0 = 0, su, sv;
                int n = 0;
                for (int dy = 0; dy < (1 << ss_y); dy++)
                    for (int dx = 0; dx < (1 << ss_x); dx++) {
                        s0 += yd[(i + dy) * bw + j + dx];
                        n++;
                    }
                su = sv = s0;
                for (int dy = -2; dy <= 2; dy++)
                    for (int dx = -2; dx <= 2; dx++) {
                        int r = ur + dy, c = uc + dx;
                        r = r < 0 ? 0 : r > (int)uh - 1 ? (int)uh - 1 : r;
                        c = c < 0 ? 0 : c > (int)uw - 1 ? (int)uw - 1 : c;
                        su += ud[r * uw + c];
                        sv += vd[r * uw + c];
                        n++;
                    }
                const unsigned m = ur * uv_pre_stride + uc;
                w = orc_tf_weight(su >> sh, n, block_error[q], d_factor[q], den[1]);
                u_count[m] += (uint16_t)w;
                u_accum[m] += (uint32_t)(w * px(u_pre, m, hbd));
                w = orc_tf_weight(sv >> sh, n, block_error[q], d_factor[q], den[2]);
                v_count[m] += (uint16_t)w;
                v_accum[m] += (uint32_t)(w * px(v_pre, m, hbd));
            }
        }
    free(yd);
}

/* apply_filtering_central (:551-621): the frame being filtered enters with weight 1000 at every sample. */
void orc_tf_central(const void *pre, int pre_stride, unsigned w, unsigned h, int hbd, uint32_t *accum, uint16_t *count, int acc_stride) {
    for (unsigned i = 0; i < h; i++)
        for (unsigned j = 0; j < w; j++) {
            accum[i * acc_stride + j] += 1000u * (uint32_t)px(pre, i * pre_stride + j, hbd);
            count[i * acc_stride + j] += 1000;
        }
}

/* get_final_filtered_pixels (:1943-2050): dst = (accum + count / 2) / count; returns the sum of (dst_before - new)^2. */
uint64_t orc_tf_normalize(void *dst, int dst_stride, unsigned w, unsigned h, int hbd, const uint32_t *accum, const uint16_t *count,
                          int acc_stride) {
    uint64_t sse = 0;
    for (unsigned i = 0; i < h; i++)
        for (unsigned j = 0; j < w; j++) {
            const uint32_t c = count[i * acc_stride + j], v = (accum[i * acc_stride + j] + (c >> 1)) / c;
            const int old = px(dst, i * dst_stride + j, hbd), d = old - (int)v;
            sse += (uint64_t)((int64_t)d * d);
            if (hbd)
                ((uint16_t *)dst)[i * dst_stride + j] = (uint16_t)v;
            else
                ((uint8_t *)dst)[i * dst_stride + j] = (uint8_t)v;
        }
    return sse;
}
