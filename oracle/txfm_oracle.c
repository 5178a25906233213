/*
 * txfm_oracle.c — CPU restatement (plain scalar C) of the SVT-AV1 v0.8.6 integer transforms:
 * forward / inverse 1-D DCT 4..64, ADST 4/8/16, identity 4..64, the separable 2-D cores, the 64-wide
 * zero-out/re-pack helpers, residual and the quantisers.
 *
 * TEST INFRASTRUCTURE ONLY (see me_oracle.c).  Pinned bit-for-bit against the reference's own C functions in
 * oracle/_ref (svt_av1_fdct*_new, svt_av1_idct*_new, svt_av1_f/iadst*_new, the 19 svt_av1_fwd_txfm2d_* and
 * svt_av1_inv_txfm2d_add_* entries, svt_aom_quantize_b, svt_av1_quantize_fp ...) — tests/test_oracle_txfm.py.
 *
 * The reference spells every butterfly stage out by hand (Encoder/Codec/EbTransforms.c:75-2271,
 * Common/Codec/EbInvTransforms.c:75-2358).  Here each network is generated from its structure instead:
 *   DCT-n  = butterfly(n) ; DCT-(n/2) on the sums ; ODD(n/2) on the differences ; bit-reversal
 *   ODD(m) = alternating "rotation" (R_j) and butterfly (B_j) layers and a final rotation layer
 * Only half_btf() rounds, additions are exact, and every layer works on disjoint pairs, so the order in which
 * independent sub-networks run does not change a single bit; the inverse is the same layers transposed and
 * reversed, with the reference's clamp after every add/sub.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/svt_av1_b200.h"
#include "oracle.h"

/* ---- constants ------------------------------------------------------------------------------------- */
static int32_t g_cospi[4][64]; /* cos_bit 10..13: round(cos(i*pi/128) * 2^bit), eb_av1_cospi_arr_data */
static int32_t g_sinpi[4][5]; /* round(sin(i*pi/9) * 2^bit * 2*sqrt(2)/3), eb_av1_sinpi_arr_data */
static int g_tables_ready = 0;

static void init_tables(void) {
    if (g_tables_ready) return;
    for (int b = 10; b <= 13; b++) {
        for (int i = 0; i < 64; i++) g_cospi[b - 10][i] = (int32_t)floor(cos(M_PI * i / 128.0) * (double)(1 << b) + 0.5);
        /* sin(i*pi/9) * 2^bit * 2*sqrt(2)/3, hand-adjusted in the reference so that sinpi[1]+sinpi[2]==sinpi[4]
         * (eb_av1_sinpi_arr_data, Common/Codec/EbInvTransforms.c): not derivable by rounding, kept as data */
        static const int32_t k_sinpi[4][5] = {{0, 330, 621, 836, 951}, {0, 660, 1241, 1672, 1901},
                                              {0, 1321, 2482, 3344, 3803}, {0, 2642, 4964, 6689, 7606}};
        for (int i = 0; i < 5; i++) g_sinpi[b - 10][i] = k_sinpi[b - 10][i];
    }
    g_tables_ready = 1;
}
ORC_API const int32_t *orc_cospi_table(int bit) {
    init_tables();
    return g_cospi[bit - 10];
}
ORC_API const int32_t *orc_sinpi_table(int bit) {
    init_tables();
    return g_sinpi[bit - 10];
}

/* Common/Codec/EbInvTransforms.h:285-312 */
static inline int32_t round_shift64(int64_t v, int bit) { return (int32_t)((v + ((int64_t)1 << (bit - 1))) >> bit); }
static inline int32_t half_btf(int32_t w0, int32_t in0, int32_t w1, int32_t in1, int bit) {
    /* products wrap in 32 bits exactly as `(int64_t)(w0 * in0)` does in the reference */
    const int64_t r = (int64_t)(int32_t)((uint32_t)w0 * (uint32_t)in0) + (int64_t)(int32_t)((uint32_t)w1 * (uint32_t)in1);
    return (int32_t)((r + ((int64_t)1 << (bit - 1))) >> bit);
}
static inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
static inline int32_t wsub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
/* clamp_value, EbInvTransforms.c:67-73 */
static inline int32_t clampv(int32_t v, int bit) {
    if (bit <= 0) return v;
    const int64_t mx = ((int64_t)1 << (bit - 1)) - 1, mn = -((int64_t)1 << (bit - 1));
    return (int32_t)(v < mn ? mn : (v > mx ? mx : v));
}
static int ilog2(int n) {
    int l = 0;
    while ((1 << l) < n) l++;
    return l;
}
static int brev(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; i++)
        if (v & (1 << i)) r |= 1 << (bits - 1 - i);
    return r;
}

/* ---- DCT ------------------------------------------------------------------------------------------- */
/* x: strided array (element i at x[i*s]). All layers act on disjoint pairs, so they run in place. */
#define X(i) x[(size_t)(i) * s]

/* butterfly layer of span S inside y[0..m): block q even -> "normal", odd -> "mirrored" */
static void odd_butterflies(int32_t *x, int s, int m, int S, int clamp_bit) {
    for (int base = 0; base < m; base += S) {
        const int mirrored = (base / S) & 1;
        for (int i = 0; i < S / 2; i++) {
            const int lo = base + i, hi = base + S - 1 - i;
            const int32_t a = X(lo), b = X(hi);
            int32_t sum = wadd(a, b), dif = mirrored ? wsub(b, a) : wsub(a, b);
            if (clamp_bit) {
                sum = clampv(sum, clamp_bit);
                dif = clampv(dif, clamp_bit);
            }
            if (mirrored) {
                X(hi) = sum;
                X(lo) = dif;
            } else {
                X(lo) = sum;
                X(hi) = dif;
            }
        }
    }
}
/* rotation layer R_j of ODD(m): pairs (t, m-1-t); symmetric matrices, identical in both directions */
static void odd_rotations(int32_t *x, int s, int m, int j, const int32_t *c, int bit) {
    const int G = m >> j;
    for (int t = 0; t < m / 2; t++) {
        const int u = t % G, p = m - 1 - t;
        const int k = (32 >> j) * brev((1 << j) + t / G, j + 1);
        const int32_t a = X(t), b = X(p);
        if (u >= G / 4 && u < G / 2) { /* type A */
            X(t) = half_btf(-c[k], a, c[64 - k], b, bit);
            X(p) = half_btf(c[k], b, c[64 - k], a, bit);
        } else if (u >= G / 2 && u < 3 * G / 4) { /* type B */
            X(t) = half_btf(-c[64 - k], a, -c[k], b, bit);
            X(p) = half_btf(c[64 - k], b, -c[k], a, bit);
        }
    }
}
/* last (forward) / first (inverse) layer of ODD(m) inside a DCT-n (n = 2m): angles in bit-reversed order */
static void odd_final_rotation(int32_t *x, int s, int m, const int32_t *c, int bit, int inverse) {
    const int n = 2 * m, L = ilog2(n);
    for (int t = 0; t < m / 2; t++) {
        const int p = m - 1 - t, k = (64 / n) * brev(m + t, L);
        const int32_t a = X(t), b = X(p);
        if (!inverse) {
            X(t) = half_btf(c[64 - k], a, c[k], b, bit);
            X(p) = half_btf(c[64 - k], b, -c[k], a, bit);
        } else {
            X(t) = half_btf(c[64 - k], a, -c[k], b, bit);
            X(p) = half_btf(c[k], a, c[64 - k], b, bit);
        }
    }
}

void orc_fdct(int32_t *x, int s, int n_total, int bit) {
    init_tables();
    const int32_t *c = g_cospi[bit - 10];
    for (int n = n_total; n >= 4; n >>= 1) {
        const int m = n / 2, L = ilog2(n);
        for (int i = 0; i < m; i++) { /* stage-1 butterfly of DCT-n */
            const int32_t a = X(i), b = X(n - 1 - i);
            X(i) = wadd(a, b);
            X(n - 1 - i) = wsub(a, b);
        }
        int32_t *y = x + (size_t)m * s; /* ODD(m) on the differences */
        for (int j = 0; j <= L - 3; j++) {
            odd_rotations(y, s, m, j, c, bit);
            odd_butterflies(y, s, m, m >> (j + 1), 0);
        }
        odd_final_rotation(y, s, m, c, bit, 0);
    }
    { /* DCT-2 */
        const int32_t a = X(0), b = X(1);
        X(0) = half_btf(c[32], a, c[32], b, bit);
        X(1) = half_btf(-c[32], b, c[32], a, bit);
    }
    const int L = ilog2(n_total);
    for (int j = 0; j < n_total; j++) {
        const int r = brev(j, L);
        if (r > j) {
            const int32_t t = X(j);
            X(j) = X(r);
            X(r) = t;
        }
    }
}

void orc_idct(int32_t *x, int s, int n_total, int bit, int clamp_bit) {
    init_tables();
    const int32_t *c = g_cospi[bit - 10];
    const int Lt = ilog2(n_total);
    for (int j = 0; j < n_total; j++) {
        const int r = brev(j, Lt);
        if (r > j) {
            const int32_t t = X(j);
            X(j) = X(r);
            X(r) = t;
        }
    }
    {
        const int32_t a = X(0), b = X(1);
        X(0) = half_btf(c[32], a, c[32], b, bit);
        X(1) = half_btf(c[32], a, -c[32], b, bit);
    }
    for (int n = 4; n <= n_total; n <<= 1) {
        const int m = n / 2, L = ilog2(n);
        int32_t *y = x + (size_t)m * s;
        odd_final_rotation(y, s, m, c, bit, 1);
        for (int j = L - 3; j >= 0; j--) {
            odd_butterflies(y, s, m, m >> (j + 1), clamp_bit);
            odd_rotations(y, s, m, j, c, bit);
        }
        for (int i = 0; i < m; i++) {
            const int32_t a = X(i), b = X(n - 1 - i);
            X(i) = clampv(wadd(a, b), clamp_bit);
            X(n - 1 - i) = clampv(wsub(a, b), clamp_bit);
        }
    }
}

/* ---- ADST ------------------------------------------------------------------------------------------ */
/* fadst4 / iadst4: Encoder/Codec/EbTransforms.c:1445-1533, Common/Codec/EbInvTransforms.c:707-792 */
void orc_fadst4(int32_t *x, int s, int bit) {
    init_tables();
    const int32_t *sp = g_sinpi[bit - 10];
    const int32_t x0 = X(0), x1 = X(1), x2 = X(2), x3 = X(3);
    if (!(x0 | x1 | x2 | x3)) {
        X(0) = X(1) = X(2) = X(3) = 0;
        return;
    }
#define M(a, b) ((int32_t)((uint32_t)(a) * (uint32_t)(b)))
    const int32_t s0 = M(sp[1], x0), s1 = M(sp[4], x0), s2 = M(sp[2], x1), s3 = M(sp[1], x1), s4 = M(sp[3], x2),
                  s5 = M(sp[4], x3), s6 = M(sp[2], x3), s7 = wsub(wadd(x0, x1), x3);
    const int32_t y0 = wadd(wadd(s0, s2), s5), y1 = M(sp[3], s7), y2 = wadd(wsub(s1, s3), s6), y3 = s4;
    X(0) = round_shift64((int64_t)wadd(y0, y3), bit);
    X(1) = round_shift64((int64_t)y1, bit);
    X(2) = round_shift64((int64_t)wsub(y2, y3), bit);
    X(3) = round_shift64((int64_t)wadd(wsub(y2, y0), y3), bit);
}
void orc_iadst4(int32_t *x, int s, int bit) {
    init_tables();
    const int32_t *sp = g_sinpi[bit - 10];
    const int32_t x0 = X(0), x1 = X(1), x2 = X(2), x3 = X(3);
    if (!(x0 | x1 | x2 | x3)) {
        X(0) = X(1) = X(2) = X(3) = 0;
        return;
    }
    /* EbInvTransforms.c:742-790: products and sums in (wrapping) int32, rounding in int64 */
    int32_t s0 = M(sp[1], x0), s1 = M(sp[2], x0), s2 = M(sp[3], x1), s3 = M(sp[4], x2), s4 = M(sp[1], x2),
            s5 = M(sp[2], x3), s6 = M(sp[4], x3), s7 = wadd(wsub(x0, x2), x3);
    s0 = wadd(wadd(s0, s3), s5);
    s1 = wsub(wsub(s1, s4), s6);
    s3 = s2;
    s2 = M(sp[3], s7);
    X(0) = round_shift64((int64_t)wadd(s0, s3), bit);
    X(1) = round_shift64((int64_t)wadd(s1, s3), bit);
    X(2) = round_shift64((int64_t)s2, bit);
    X(3) = round_shift64((int64_t)wsub(wadd(s0, s1), s3), bit);
#undef M
}

/* ADST-8/16 (EbTransforms.c:1535-1826, EbInvTransforms.c:797-1105): signed input permutation, then for span
 * h = 2,4,..,n/2 a rotation layer on the upper half of every group of 2h followed by an add/sub layer of stride h,
 * a final rotation layer on adjacent pairs and an output permutation.  The inverse runs the same layers backwards
 * (each rotation matrix is symmetric or its own transpose partner) with clamps after add/sub. */
static const int8_t k_adst8_in[8] = {0, -7, -3, 4, -1, 6, 2, -5};
static const int8_t k_adst8_out[8] = {1, 6, 3, 4, 5, 2, 7, 0};
static const int8_t k_adst16_in[16] = {0, -15, -7, 8, -3, 12, 4, -11, -1, 14, 6, -9, 2, -13, -5, 10};
static const int8_t k_adst16_out[16] = {1, 14, 3, 12, 5, 10, 7, 8, 9, 6, 11, 4, 13, 2, 15, 0};

/* rotation layer for span h (h >= 2): inside each group of 2h the pairs (2i, 2i+1) of the upper half */
static void adst_rotations(int32_t *v, int n, int h, const int32_t *c, int bit) {
    const int np = h / 2; /* pairs in the upper half */
    for (int g = 0; g < n; g += 2 * h)
        for (int i = 0; i < np; i++) {
            const int a0 = g + h + 2 * i, a1 = a0 + 1;
            const int32_t a = v[a0], b = v[a1];
            if (h == 2) { /* c32 pair */
                v[a0] = half_btf(c[32], a, c[32], b, bit);
                v[a1] = half_btf(c[32], a, -c[32], b, bit);
                continue;
            }
            const int half = np / 2, q = i % half;
            const int k = (64 / h) * (h >= 8 ? 4 * q + 1 : 1); /* h=4: 16; h=8: 8, 40 */
            if (i < half) {
                v[a0] = half_btf(c[k], a, c[64 - k], b, bit);
                v[a1] = half_btf(c[64 - k], a, -c[k], b, bit);
            } else {
                v[a0] = half_btf(-c[64 - k], a, c[k], b, bit);
                v[a1] = half_btf(c[k], a, c[64 - k], b, bit);
            }
        }
}
static void adst_addsub(int32_t *v, int n, int h, int clamp_bit) {
    for (int g = 0; g < n; g += 2 * h)
        for (int i = 0; i < h; i++) {
            const int32_t a = v[g + i], b = v[g + h + i];
            v[g + i] = clamp_bit ? clampv(wadd(a, b), clamp_bit) : wadd(a, b);
            v[g + h + i] = clamp_bit ? clampv(wsub(a, b), clamp_bit) : wsub(a, b);
        }
}
static void adst_final(int32_t *v, int n, const int32_t *c, int bit) {
    for (int i = 0; i < n / 2; i++) {
        const int k = (32 + 128 * i) / n;
        const int32_t a = v[2 * i], b = v[2 * i + 1];
        v[2 * i] = half_btf(c[k], a, c[64 - k], b, bit);
        v[2 * i + 1] = half_btf(c[64 - k], a, -c[k], b, bit);
    }
}
void orc_fadst(int32_t *x, int s, int n, int bit) {
    if (n == 4) {
        orc_fadst4(x, s, bit);
        return;
    }
    init_tables();
    const int32_t *c = g_cospi[bit - 10];
    const int8_t *pin = n == 8 ? k_adst8_in : k_adst16_in, *pout = n == 8 ? k_adst8_out : k_adst16_out;
    int32_t v[16], w[16];
    for (int i = 0; i < n; i++) {
        const int src = pin[i] < 0 ? -pin[i] : pin[i];
        v[i] = (pin[i] < 0) ? wsub(0, X(src)) : X(src);
    }
    for (int h = 2; h < n; h <<= 1) {
        adst_rotations(v, n, h, c, bit);
        adst_addsub(v, n, h, 0);
    }
    adst_final(v, n, c, bit);
    for (int i = 0; i < n; i++) w[i] = v[pout[i]];
    for (int i = 0; i < n; i++) X(i) = w[i];
}
void orc_iadst(int32_t *x, int s, int n, int bit, int clamp_bit) {
    if (n == 4) {
        orc_iadst4(x, s, bit);
        return;
    }
    init_tables();
    const int32_t *c = g_cospi[bit - 10];
    const int8_t *pin = n == 8 ? k_adst8_in : k_adst16_in, *pout = n == 8 ? k_adst8_out : k_adst16_out;
    int32_t v[16], w[16];
    /* the inverse starts from the forward's output permutation read backwards ... */
    for (int i = 0; i < n; i++) v[pout[i]] = X(i);
    adst_final(v, n, c, bit); /* each 2x2 of the final layer is symmetric */
    for (int h = n / 2; h >= 2; h >>= 1) {
        adst_addsub(v, n, h, clamp_bit);
        adst_rotations(v, n, h, c, bit);
    }
    /* ... and ends with the forward's signed input permutation read backwards */
    for (int i = 0; i < n; i++) {
        const int dst = pin[i] < 0 ? -pin[i] : pin[i];
        w[dst] = (pin[i] < 0) ? wsub(0, v[i]) : v[i];
    }
    for (int i = 0; i < n; i++) X(i) = w[i];
}

/* ---- identity (EbTransforms.c:2239-2278, EbInvTransforms.c:2321-2358) ------------------------------ */
#define SQRT2 5793
void orc_fidentity(int32_t *x, int s, int n) {
    for (int i = 0; i < n; i++) {
        const int32_t v = X(i);
        if (n == 4) X(i) = round_shift64((int64_t)v * SQRT2, 12);
        else if (n == 8) X(i) = (int32_t)((uint32_t)v * 2u);
        else if (n == 16) X(i) = round_shift64((int64_t)v * 2 * SQRT2, 12);
        else if (n == 32) X(i) = (int32_t)((uint32_t)v * 4u);
        else X(i) = round_shift64((int64_t)v * 4 * SQRT2, 12);
    }
}
#undef X

/* ---- 2-D ------------------------------------------------------------------------------------------- */
static const uint8_t k_txw[19] = {4, 8, 16, 32, 64, 4, 8, 8, 16, 16, 32, 32, 64, 4, 16, 8, 32, 16, 64};
static const uint8_t k_txh[19] = {4, 8, 16, 32, 64, 8, 4, 16, 8, 32, 16, 64, 32, 16, 4, 32, 8, 64, 16};
/* fwd_txfm_shift_ls (Encoder/Codec/EbTransforms.h:26-44) */
static const int8_t k_fwd_shift[19][3] = {{2, 0, 0},  {2, -1, 0},  {2, -2, 0},  {2, -4, 0}, {0, -2, -2}, {2, -1, 0}, {2, -1, 0},
                                          {2, -2, 0}, {2, -2, 0},  {2, -4, 0},  {2, -4, 0}, {0, -2, -2}, {2, -4, -2}, {2, -1, 0},
                                          {2, -1, 0}, {2, -2, 0},  {2, -2, 0},  {0, -2, 0}, {2, -4, 0}};
/* eb_inv_txfm_shift_ls (Common/Codec/EbInvTransforms.h:51-69) */
static const int8_t k_inv_shift[19][2] = {{0, -4},  {-1, -4}, {-2, -4}, {-2, -4}, {-2, -4}, {0, -4},  {0, -4},  {-1, -4}, {-1, -4}, {-1, -4},
                                          {-1, -4}, {-1, -4}, {-1, -4}, {-1, -4}, {-1, -4}, {-2, -4}, {-2, -4}, {-2, -4}, {-2, -4}};
/* fwd_cos_bit_col / fwd_cos_bit_row [txw_idx][txh_idx] (Encoder/Codec/EbTransforms.h:46-57) */
static const int8_t k_fwd_cos_col[5][5] = {{13, 13, 13, 0, 0}, {13, 13, 13, 12, 0}, {13, 13, 13, 12, 13}, {0, 13, 13, 12, 13}, {0, 0, 13, 12, 13}};
static const int8_t k_fwd_cos_row[5][5] = {{13, 13, 12, 0, 0}, {13, 13, 13, 12, 0}, {13, 13, 12, 13, 12}, {0, 12, 13, 12, 11}, {0, 0, 12, 11, 10}};
/* 1-D kinds per 2-D TxType: 0 DCT, 1 ADST, 2 FLIPADST, 3 IDTX (vtx_tab / htx_tab, EbInvTransforms.h) */
static const uint8_t k_vtx[16] = {0, 1, 0, 1, 2, 0, 2, 1, 2, 3, 0, 3, 1, 3, 2, 3};
static const uint8_t k_htx[16] = {0, 0, 1, 1, 0, 2, 2, 2, 1, 3, 3, 0, 3, 1, 3, 2};

static void fwd_1d(int32_t *x, int s, int n, int kind, int bit) {
    if (kind == 0) orc_fdct(x, s, n, bit);
    else if (kind == 3) orc_fidentity(x, s, n);
    else orc_fadst(x, s, n, bit);
}
static void round_shift_arr(int32_t *x, int s, int n, int bit) { /* svt_av1_round_shift_array_c */
    if (bit == 0) return;
    for (int i = 0; i < n; i++) {
        if (bit > 0) x[(size_t)i * s] = round_shift64((int64_t)x[(size_t)i * s], bit);
        else x[(size_t)i * s] = (int32_t)((uint32_t)x[(size_t)i * s] * (1u << (-bit)));
    }
}

/* av1_tranform_two_d_core_c (EbTransforms.c:2301-2370) for any of the 19 sizes x 16 types */
void orc_fwd_txfm2d(const int16_t *input, int32_t *output, uint32_t stride, int tx_type, int tx_size, int bit_depth) {
    (void)bit_depth;
    const int w = k_txw[tx_size], h = k_txh[tx_size];
    const int8_t *shift = k_fwd_shift[tx_size];
    const int cbc = k_fwd_cos_col[ilog2(w) - 2][ilog2(h) - 2], cbr = k_fwd_cos_row[ilog2(w) - 2][ilog2(h) - 2];
    const int vk = k_vtx[tx_type], hk = k_htx[tx_type];
    const int ud = vk == 2, lr = hk == 2;
    int32_t *buf = malloc(sizeof(int32_t) * w * h);
    for (int c = 0; c < w; c++) {
        int32_t col[64];
        for (int r = 0; r < h; r++) col[r] = input[(size_t)(ud ? h - 1 - r : r) * stride + c];
        round_shift_arr(col, 1, h, -shift[0]);
        fwd_1d(col, 1, h, vk, cbc);
        round_shift_arr(col, 1, h, -shift[1]);
        for (int r = 0; r < h; r++) buf[r * w + (lr ? w - 1 - c : c)] = col[r];
    }
    const int rect = (w == 2 * h || h == 2 * w);
    for (int r = 0; r < h; r++) {
        int32_t *row = output + (size_t)r * w;
        memcpy(row, buf + r * w, sizeof(int32_t) * w);
        fwd_1d(row, 1, w, hk, cbr);
        round_shift_arr(row, 1, w, -shift[2]);
        if (rect)
            for (int c = 0; c < w; c++) row[c] = round_shift64((int64_t)row[c] * SQRT2, 12);
    }
    free(buf);
}

/* Partial-frequency variants (N2 / N4 shapes, EbTransforms.c:5401-5648, 7007-7250): the reference's N2/N4 cores
 * compute exactly the top-left (w>>s) x (h>>s) coefficients of the full transform and zero the rest. */
void orc_fwd_txfm2d_pf(const int16_t *input, int32_t *output, uint32_t stride, int tx_type, int tx_size, int bit_depth,
                       int shift) {
    orc_fwd_txfm2d(input, output, stride, tx_type, tx_size, bit_depth);
    const int w = k_txw[tx_size], h = k_txh[tx_size];
    for (int i = 0; i < w * h; i++)
        if (i % w >= (w >> shift) || i / w >= (h >> shift)) output[i] = 0;
}

/* svt_handle_transform64x64/64x32/32x64/64x16/16x64 (EbTransforms.c:2763-2931): energy of the discarded
 * area, zero it, and re-pack the kept 32-wide coefficients contiguously. Returns the energy. */
uint64_t orc_handle_transform64(int32_t *output, int tx_size) {
    const int w = k_txw[tx_size], h = k_txh[tx_size];
    const int kw = w > 32 ? 32 : w, kh = h > 32 ? 32 : h;
    uint64_t e = 0;
    for (int r = 0; r < h; r++)
        for (int c = 0; c < w; c++)
            if (r >= kh || c >= kw) {
                e += (uint64_t)((int64_t)output[r * w + c] * (int64_t)output[r * w + c]);
                output[r * w + c] = 0;
            }
    if (kw != w)
        for (int r = 1; r < kh; r++) memmove(output + r * kw, output + r * w, sizeof(int32_t) * kw);
    return e;
}

/* av1_estimate_transform (EbTransforms.c:3613-3670): the trans_coeff_shape dispatcher of the encoder.
 *   DEFAULT_SHAPE  (:3434-3600) forward transform, then for 64-sized units the discarded-area energy / zero / re-pack;
 *   N2 / N4 SHAPE  (:3052-3405) the top-left (w>>s) x (h>>s) coefficients only; 64-sized units are re-packed 32 wide by
 *                  handle_transform*_N2_N4 (:2933-2965), which report an energy of 0;
 *   ONLY_DC_SHAPE  (:3407-3430) the N4 result with every entry but coefficient 0 cleared.
 * coeff: w*h entries (the packed layout of 64-sized units occupies the first min(w,32)*min(h,32)). Returns the energy. */
uint64_t orc_estimate_transform(const int16_t *res, uint32_t stride, int32_t *coeff, int tx_size, int bit_depth, int tx_type,
                                int shape) {
    const int w = k_txw[tx_size], h = k_txh[tx_size];
    if (shape == 0) {
        orc_fwd_txfm2d(res, coeff, stride, tx_type, tx_size, bit_depth);
        return (w == 64 || h == 64) ? orc_handle_transform64(coeff, tx_size) : 0;
    }
    orc_fwd_txfm2d_pf(res, coeff, stride, tx_type, tx_size, bit_depth, shape == 1 ? 1 : 2);
    if (w == 64 || h == 64) orc_handle_transform64(coeff, tx_size); /* the dropped area is already zero: re-pack only */
    if (shape == 3) {
        /* the clearing loop of :3424-3429 walks w-wide indices over the N4 result; every entry the N4 transform left
         * non-zero satisfies its condition, so exactly coefficient 0 survives */
        for (int i = 1; i < w * h; i++)
            if (i % w < (w >> 2) || i / w < (h >> 2)) coeff[i] = 0;
    }
    return 0;
}

static void inv_1d(int32_t *x, int s, int n, int kind, int bit, int clamp_bit) {
    if (kind == 0) orc_idct(x, s, n, bit, clamp_bit);
    else if (kind == 3) { /* iidentity: same scaling as forward */
        orc_fidentity(x, s, n);
    } else orc_iadst(x, s, n, bit, clamp_bit);
}

/* inv_txfm2d_add_c (EbInvTransforms.c:2455-2532) incl. the 64-wide input expansion (:2573-2589, 2648-2714).
 * input: w'xh' coefficients where dimensions of 64 are stored as 32 (zero high half). */
void orc_inv_txfm2d_add(const int32_t *input, const uint16_t *pred, int32_t stride_r, uint16_t *recon, int32_t stride_w,
                        int tx_type, int tx_size, int bd) {
    const int w = k_txw[tx_size], h = k_txh[tx_size];
    const int iw = w > 32 ? 32 : w, ih = h > 32 ? 32 : h;
    const int8_t *shift = k_inv_shift[tx_size];
    const int vk = k_vtx[tx_type], hk = k_htx[tx_type];
    const int ud = vk == 2, lr = hk == 2;
    const int range_row = bd == 8 ? 16 : bd == 10 ? 18 : 20, range_col = bd == 8 ? 16 : bd == 10 ? 16 : 18;
    const int rect = (w == 2 * h || h == 2 * w);
    int32_t *buf = calloc((size_t)w * h, sizeof(int32_t));
    for (int r = 0; r < h; r++) {
        int32_t *row = buf + (size_t)r * w;
        for (int c = 0; c < w; c++) {
            int32_t v = (r < ih && c < iw) ? input[r * iw + c] : 0;
            if (rect) v = round_shift64((int64_t)v * 2896, 12);
            row[c] = clampv(v, bd + 8);
        }
        inv_1d(row, 1, w, hk, 12, range_row);
        round_shift_arr(row, 1, w, -shift[0]);
    }
    const int col_clamp = bd + 6 > 16 ? bd + 6 : 16;
    for (int c = 0; c < w; c++) {
        int32_t col[64];
        for (int r = 0; r < h; r++) col[r] = clampv(buf[r * w + (lr ? w - 1 - c : c)], col_clamp);
        inv_1d(col, 1, h, vk, 12, range_col);
        round_shift_arr(col, 1, h, -shift[1]);
        for (int r = 0; r < h; r++) {
            const int64_t v = (int64_t)pred[(size_t)r * stride_r + c] + col[ud ? h - 1 - r : r];
            const int mx = (1 << bd) - 1;
            recon[(size_t)r * stride_w + c] = (uint16_t)(v < 0 ? 0 : (v > mx ? mx : v));
        }
    }
    free(buf);
}

/* ---- residual (Common/Codec/EbPictureOperators.c:106-150) ----------------------------------------- */
void orc_residual(const void *src, uint32_t src_stride, const void *pred, uint32_t pred_stride, int16_t *res,
                  uint32_t res_stride, uint32_t w, uint32_t h, int hbd) {
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            const int s = hbd ? ((const uint16_t *)src)[y * src_stride + x] : ((const uint8_t *)src)[y * src_stride + x];
            const int p = hbd ? ((const uint16_t *)pred)[y * pred_stride + x] : ((const uint8_t *)pred)[y * pred_stride + x];
            res[y * res_stride + x] = (int16_t)(s - p);
        }
}

/* ---- quantisers (Encoder/Codec/EbFullLoop.c:37-93, 171-225, 314-600) ------------------------------- */
#define QM_BITS 5
static inline int rpot(int v, int n) { return n ? (v + (1 << (n - 1))) >> n : v; } /* ROUND_POWER_OF_TWO */

/* svt_aom_quantize_b_c_ii (lowbd, hbd = 0) and svt_aom_highbd_quantize_b_c (hbd = 1).  The reference's pre-scan
 * only skips coefficients the main test rejects anyway, so one pass in scan order is equivalent. */
void orc_quantize_b(const int32_t *coeff, intptr_t n, const int16_t *zbin, const int16_t *round, const int16_t *quant,
                    const int16_t *quant_shift, int32_t *qcoeff, int32_t *dqcoeff, const int16_t *dequant,
                    uint16_t *eob_ptr, const int16_t *scan, const uint8_t *qm, const uint8_t *iqm, int log_scale,
                    int hbd) {
    const int zbins[2] = {rpot(zbin[0], log_scale), rpot(zbin[1], log_scale)};
    int eob = -1;
    memset(qcoeff, 0, n * sizeof(int32_t));
    memset(dqcoeff, 0, n * sizeof(int32_t));
    for (int i = 0; i < n; i++) {
        const int rc = scan[i], ac = rc != 0;
        const int c = coeff[rc], sign = c < 0 ? -1 : 0;
        const int abs_c = (c ^ sign) - sign;
        const int wt = qm ? qm[rc] : (1 << QM_BITS), iwt = iqm ? iqm[rc] : (1 << QM_BITS);
        int keep;
        if (hbd) {
            const int cw = c * wt;
            keep = cw >= zbins[ac] * (1 << QM_BITS) || cw <= -zbins[ac] * (1 << QM_BITS);
        } else {
            keep = abs_c * wt >= (zbins[ac] << QM_BITS);
        }
        if (!keep) continue;
        int64_t tmp = (int64_t)abs_c + rpot(round[ac], log_scale);
        if (!hbd) tmp = tmp < INT16_MIN ? INT16_MIN : (tmp > INT16_MAX ? INT16_MAX : tmp);
        tmp *= wt;
        const int32_t q = (int32_t)(((((tmp * quant[ac]) >> 16) + tmp) * quant_shift[ac]) >> (16 - log_scale + QM_BITS));
        qcoeff[rc] = (q ^ sign) - sign;
        const int dq = (dequant[ac] * iwt + (1 << (QM_BITS - 1))) >> QM_BITS;
        const int32_t adq = (int32_t)((uint32_t)q * (uint32_t)dq) >> log_scale;
        dqcoeff[rc] = (adq ^ sign) - sign;
        if (q) eob = i;
    }
    *eob_ptr = (uint16_t)(eob + 1);
}

/* quantize_fp_helper_c (lowbd: svt_av1_quantize_fp / _32x32 / _64x64 with log_scale 0/1/2) and
 * highbd_quantize_fp_helper_c (svt_av1_highbd_quantize_fp), no quantisation matrices */
void orc_quantize_fp(const int32_t *coeff, intptr_t n, const int16_t *round, const int16_t *quant, int32_t *qcoeff,
                     int32_t *dqcoeff, const int16_t *dequant, uint16_t *eob_ptr, const int16_t *scan, int log_scale,
                     int hbd) {
    const int rounding[2] = {rpot(round[0], log_scale), rpot(round[1], log_scale)};
    int eob = -1;
    memset(qcoeff, 0, n * sizeof(int32_t));
    memset(dqcoeff, 0, n * sizeof(int32_t));
    for (int i = 0; i < n; i++) {
        const int rc = scan[i], ac = rc != 0;
        const int c = coeff[rc], sign = c < 0 ? -1 : 0;
        int64_t abs_c = (c ^ sign) - sign;
        int q = 0;
        if (hbd) {
            if (((int)abs_c << (1 + log_scale)) >= dequant[ac]) {
                q = (int)(((abs_c + rounding[ac]) * quant[ac]) >> (16 - log_scale));
                qcoeff[rc] = (q ^ sign) - sign;
                const int32_t adq = (int32_t)((uint32_t)q * (uint32_t)dequant[ac]) >> log_scale;
                dqcoeff[rc] = (adq ^ sign) - sign;
            }
        } else if ((abs_c << (1 + log_scale)) >= (int32_t)dequant[ac]) {
            abs_c += rounding[ac];
            abs_c = abs_c < INT16_MIN ? INT16_MIN : (abs_c > INT16_MAX ? INT16_MAX : abs_c);
            q = (int)((abs_c * quant[ac]) >> (16 - log_scale));
            if (q) {
                qcoeff[rc] = (q ^ sign) - sign;
                const int32_t adq = (int32_t)((uint32_t)q * (uint32_t)dequant[ac]) >> log_scale;
                dqcoeff[rc] = (adq ^ sign) - sign;
            }
        }
        if (q) eob = i;
    }
    *eob_ptr = (uint16_t)(eob + 1);
}
