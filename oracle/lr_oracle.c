/*
 * lr_oracle.c — CPU restatement (plain scalar C) of the SVT-AV1 v0.8.6 loop-restoration kernels: the self-guided
 * filter (box sums -> a/b -> weighted 3x3), its projection apply, and the separable Wiener filter.
 *
 * TEST INFRASTRUCTURE ONLY (see me_oracle.c).  Pinned bit-for-bit against svt_av1_selfguided_restoration_c,
 * svt_apply_selfguided_restoration_c, svt_av1_wiener_convolve_add_src_c and svt_av1_highbd_wiener_convolve_add_src_c
 * of the reference compiled into oracle/_ref — tests/test_oracle_lr.py.
 * Reference paths are relative to /root/reference/Source/Lib/Common/Codec.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/svt_av1_b200.h"
#include "oracle.h"

/* eb_sgr_params (EbRestoration.c:136-153): {r0, r1}, {s0, s1} */
static const int k_sgr_r[16][2] = {{2, 1}, {2, 1}, {2, 1}, {2, 1}, {2, 1}, {2, 1}, {2, 1}, {2, 1},
                                   {2, 1}, {2, 1}, {0, 1}, {0, 1}, {0, 1}, {0, 1}, {2, 0}, {2, 0}};
static const int k_sgr_s[16][2] = {{140, 3236}, {112, 2158}, {93, 1618}, {80, 1438}, {70, 1295}, {58, 1177}, {47, 1079}, {37, 996},
                                   {30, 925},   {25, 863},   {-1, 2589}, {-1, 1618}, {-1, 1177}, {-1, 925},  {56, -1},   {22, -1}};
static int x_by_xplus1(int z) { /* eb_x_by_xplus1 (:720-736): round(256 z/(z+1)), with 0 -> 1 and 255 -> 256 */
    if (z == 0) return 1;
    if (z >= 255) return 256;
    return (256 * z + (z + 1) / 2) / (z + 1);
}
static int one_by_x(int n) { return (4096 + n / 2) / n; } /* eb_one_by_x (:738-741) */
#define RPOT(v, n) (((n) != 0) ? (((v) + (1u << ((n)-1))) >> (n)) : (v))

/* selfguided_restoration_[fast_]internal (:744-1010). dgd: int32 samples with a 3-sample border, stride ds. */
static void sgr_pass(const int32_t *dgd, int w, int h, int ds, int32_t *dst, int dst_stride, int bd, int idx, int pass) {
    const int r = k_sgr_r[idx][pass], n = (2 * r + 1) * (2 * r + 1);
    const uint32_t s = (uint32_t)k_sgr_s[idx][pass];
    const int bs = w + 2;
    int32_t *A = malloc(sizeof(int32_t) * (w + 2) * (h + 2)), *B = malloc(sizeof(int32_t) * (w + 2) * (h + 2));
    const int fast = pass == 0; /* radius index 0 is the "fast" (every other row) filter */
    for (int i = -1; i < h + 1; i += fast ? 2 : 1)
        for (int j = -1; j < w + 1; j++) {
            uint32_t sum = 0, sq = 0;
            for (int y = -r; y <= r; y++)
                for (int x = -r; x <= r; x++) {
                    const int32_t v = dgd[(i + y) * ds + j + x];
                    sum += (uint32_t)v;
                    sq += (uint32_t)(v * v);
                }
            const uint32_t a = RPOT(sq, 2 * (bd - 8)), b = RPOT(sum, bd - 8);
            const uint32_t p = (a * n < b * b) ? 0 : a * n - b * b;
            const uint32_t z = RPOT(p * s, 20);
            const int av = x_by_xplus1(z < 255 ? (int)z : 255);
            A[(i + 1) * bs + j + 1] = av;
            B[(i + 1) * bs + j + 1] = (int32_t)RPOT((uint32_t)(256 - av) * sum * (uint32_t)one_by_x(n), 12);
        }
#define AA(i, j) A[((i) + 1) * bs + (j) + 1]
#define BB(i, j) B[((i) + 1) * bs + (j) + 1]
    for (int i = 0; i < h; i++)
        for (int j = 0; j < w; j++) {
            int32_t a, b, nb;
            if (!fast) {
                nb = 5;
                a = (AA(i, j) + AA(i, j - 1) + AA(i, j + 1) + AA(i - 1, j) + AA(i + 1, j)) * 4 +
                    (AA(i - 1, j - 1) + AA(i + 1, j - 1) + AA(i - 1, j + 1) + AA(i + 1, j + 1)) * 3;
                b = (BB(i, j) + BB(i, j - 1) + BB(i, j + 1) + BB(i - 1, j) + BB(i + 1, j)) * 4 +
                    (BB(i - 1, j - 1) + BB(i + 1, j - 1) + BB(i - 1, j + 1) + BB(i + 1, j + 1)) * 3;
            } else if (!(i & 1)) {
                nb = 5;
                a = (AA(i - 1, j) + AA(i + 1, j)) * 6 + (AA(i - 1, j - 1) + AA(i + 1, j - 1) + AA(i - 1, j + 1) + AA(i + 1, j + 1)) * 5;
                b = (BB(i - 1, j) + BB(i + 1, j)) * 6 + (BB(i - 1, j - 1) + BB(i + 1, j - 1) + BB(i - 1, j + 1) + BB(i + 1, j + 1)) * 5;
            } else {
                nb = 4;
                a = AA(i, j) * 6 + (AA(i, j - 1) + AA(i, j + 1)) * 5;
                b = BB(i, j) * 6 + (BB(i, j - 1) + BB(i, j + 1)) * 5;
            }
            const int32_t v = a * dgd[i * ds + j] + b;
            const int sh = 8 + nb - 4;
            dst[i * dst_stride + j] = (v + (1 << (sh - 1))) >> sh;
        }
#undef AA
#undef BB
    free(A);
    free(B);
}

/* svt_av1_selfguided_restoration_c (:1012-1045). dgd points at sample (0,0); a 3-sample border must be readable. */
void orc_selfguided_restoration(const void *dgd, int hbd, int width, int height, int dgd_stride, int32_t *flt0,
                                int32_t *flt1, int flt_stride, int sgr_params_idx, int bit_depth) {
    const int ds = width + 6;
    int32_t *buf = malloc(sizeof(int32_t) * ds * (height + 6));
    int32_t *d = buf + 3 * ds + 3;
    for (int i = -3; i < height + 3; i++)
        for (int j = -3; j < width + 3; j++)
            d[i * ds + j] = hbd ? ((const uint16_t *)dgd)[(ptrdiff_t)i * dgd_stride + j] : ((const uint8_t *)dgd)[(ptrdiff_t)i * dgd_stride + j];
    if (k_sgr_r[sgr_params_idx][0] > 0) sgr_pass(d, width, height, ds, flt0, flt_stride, bit_depth, sgr_params_idx, 0);
    if (k_sgr_r[sgr_params_idx][1] > 0) sgr_pass(d, width, height, ds, flt1, flt_stride, bit_depth, sgr_params_idx, 1);
    free(buf);
}

/* svt_apply_selfguided_restoration_c (:1047-1084) incl. svt_decode_xq (:707-718) */
void orc_apply_selfguided_restoration(const void *dat, int hbd, int width, int height, int stride, int eps, const int32_t *xqd,
                                      void *dst, int dst_stride, int bit_depth) {
    int32_t *flt0 = malloc(sizeof(int32_t) * width * height), *flt1 = malloc(sizeof(int32_t) * width * height);
    orc_selfguided_restoration(dat, hbd, width, height, stride, flt0, flt1, width, eps, bit_depth);
    const int r0 = k_sgr_r[eps][0], r1 = k_sgr_r[eps][1];
    int32_t xq[2];
    if (r0 == 0) {
        xq[0] = 0;
        xq[1] = 128 - xqd[1];
    } else if (r1 == 0) {
        xq[0] = xqd[0];
        xq[1] = 0;
    } else {
        xq[0] = xqd[0];
        xq[1] = 128 - xq[0] - xqd[1];
    }
    const int mx = (1 << bit_depth) - 1;
    for (int i = 0; i < height; i++)
        for (int j = 0; j < width; j++) {
            const int k = i * width + j;
            const int pre = hbd ? ((const uint16_t *)dat)[(ptrdiff_t)i * stride + j] : ((const uint8_t *)dat)[(ptrdiff_t)i * stride + j];
            const int32_t u = pre << 4;
            int32_t v = u << 7;
            if (r0 > 0) v += xq[0] * (flt0[k] - u);
            if (r1 > 0) v += xq[1] * (flt1[k] - u);
            const int16_t wv = (int16_t)((v + (1 << 10)) >> 11);
            const int out = wv < 0 ? 0 : (wv > mx ? mx : wv);
            if (hbd)
                ((uint16_t *)dst)[(ptrdiff_t)i * dst_stride + j] = (uint16_t)out;
            else
                ((uint8_t *)dst)[(ptrdiff_t)i * dst_stride + j] = (uint8_t)out;
        }
    free(flt0);
    free(flt1);
}

/* svt_av1_wiener_convolve_add_src_c / svt_av1_highbd_wiener_convolve_add_src_c (convolve.c:53-147, 150-260):
 * 8-tap (7 used) separable filter with the centre sample added back; x/y steps are 16 (no scaling).
 * src points at sample (0,0); 3 samples left/right and 3 rows above / 4 below must be readable. */
void orc_wiener_convolve_add_src(const void *src, int hbd, ptrdiff_t src_stride, void *dst, ptrdiff_t dst_stride,
                                 const int16_t *filter_x, const int16_t *filter_y, int w, int h, int round_0, int round_1,
                                 int bd) {
    const int ih = h + 7;
    uint16_t *tmp = malloc(sizeof(uint16_t) * w * ih);
    const int lim = (1 << (bd + 1 + 7 - round_0)) - 1;
    for (int y = 0; y < ih; y++)
        for (int x = 0; x < w; x++) {
            int32_t sum = 0;
            for (int k = 0; k < 8; k++) {
                const ptrdiff_t o = (ptrdiff_t)(y - 3) * src_stride + x - 3 + k;
                sum += (hbd ? ((const uint16_t *)src)[o] : ((const uint8_t *)src)[o]) * filter_x[k];
            }
            const ptrdiff_t oc = (ptrdiff_t)(y - 3) * src_stride + x;
            sum += ((int32_t)(hbd ? ((const uint16_t *)src)[oc] : ((const uint8_t *)src)[oc]) << 7) + (1 << (bd + 7 - 1));
            int32_t v = (sum + (1 << (round_0 - 1))) >> round_0;
            tmp[y * w + x] = (uint16_t)(v < 0 ? 0 : (v > lim ? lim : v));
        }
    const int mx = (1 << bd) - 1;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int32_t sum = 0;
            for (int k = 0; k < 8; k++) sum += tmp[(y + k) * w + x] * filter_y[k];
            sum += ((int32_t)tmp[(y + 3) * w + x] << 7) - (1 << (bd + round_1 - 1));
            int32_t v = (sum + (1 << (round_1 - 1))) >> round_1;
            v = v < 0 ? 0 : (v > mx ? mx : v);
            if (hbd)
                ((uint16_t *)dst)[(ptrdiff_t)y * dst_stride + x] = (uint16_t)v;
            else
                ((uint8_t *)dst)[(ptrdiff_t)y * dst_stride + x] = (uint8_t)v;
        }
    free(tmp);
}
