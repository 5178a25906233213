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

/* ====================================================================================================================
 * svt_av1_loop_restoration_filter_frame (Common/Codec/EbRestoration.c:1293-1364), restated the way the reference does
 * it: a working copy of the CDEF picture extended by 3 samples (svt_extend_frame :216-260), boundary lines saved from
 * the deblocked and CDEF pictures (svt_av1_loop_restoration_save_boundary_lines :1645-1868), and per restoration unit
 * / stripe the overwrite - filter - restore protocol (svt_av1_loop_restoration_filter_unit :1162-1261,
 * setup/restore_processing_stripe_boundary :353-507).  Single tile, no super-resolution.  Samples are handled as
 * uint16 internally (the arithmetic is the same for both depths).
 * ================================================================================================================== */
#define LR_EXT 8 /* working-plane border (>= RESTORATION_BORDER 3 + the 8th Wiener tap) */
typedef struct {
    uint16_t *buf; /* (h + 2 EXT) x stride, origin at (EXT, EXT) */
    int w, h, stride;
} WorkPlane;
static uint16_t *wp_at(const WorkPlane *p, int x, int y) { return p->buf + (size_t)(y + LR_EXT) * p->stride + x + LR_EXT; }
static int plane_px(const void *base, int hbd, int stride, int x, int y) {
    return hbd ? ((const uint16_t *)base)[(size_t)y * stride + x] : ((const uint8_t *)base)[(size_t)y * stride + x];
}

void orc_lr_frame(const SvtB200LrFrameParams *p, const SvtB200Frame *cdef, const SvtB200Frame *dblk, const SvtB200Frame *out,
                  const SvtB200LrUnit *const units[3] /* host arrays */) {
    const int hbd = cdef->bit_depth > 8, bd = cdef->bit_depth;
    for (int plane = 0; plane < 3; plane++) {
        const int ss = plane ? 1 : 0;
        const int pw = (cdef->width + ss) >> ss, ph = (cdef->height + ss) >> ss;
        const void *cb = plane == 0 ? cdef->y : plane == 1 ? cdef->cb : cdef->cr;
        const void *db = plane == 0 ? dblk->y : plane == 1 ? dblk->cb : dblk->cr;
        void *ob = plane == 0 ? out->y : plane == 1 ? out->cb : out->cr;
        const int cs = plane ? cdef->stride_c : cdef->stride_y, dstr = plane ? dblk->stride_c : dblk->stride_y,
                  os = plane ? out->stride_c : out->stride_y;
        const SvtB200LrPlane *rp = &p->plane[plane];
        if (!rp->frame_restoration_type) { /* plane untouched */
            for (int y = 0; y < ph; y++)
                for (int x = 0; x < pw; x++) {
                    const int v = plane_px(cb, hbd, cs, x, y);
                    if (hbd) ((uint16_t *)ob)[(size_t)y * os + x] = (uint16_t)v;
                    else ((uint8_t *)ob)[(size_t)y * os + x] = (uint8_t)v;
                }
            continue;
        }
        /* working copy, extended by replication (only RESTORATION_BORDER = 3 is ever read through a filter tap with
         * a non-zero weight; the wider border keeps the 8th Wiener tap inside the buffer) */
        WorkPlane wk = {NULL, pw, ph, pw + 2 * LR_EXT};
        wk.buf = (uint16_t *)calloc((size_t)(ph + 2 * LR_EXT) * wk.stride, 2);
        for (int y = -3; y < ph + 3; y++)
            for (int x = -3; x < pw + 3; x++) {
                const int yy = y < 0 ? 0 : y >= ph ? ph - 1 : y, xx = x < 0 ? 0 : x >= pw ? pw - 1 : x;
                *wp_at(&wk, x, y) = (uint16_t)plane_px(cb, hbd, cs, xx, yy);
            }
        /* boundary lines: per stripe 2 rows above and 2 rows below, each extended by 4 columns */
        const int SH = 64 >> ss, off = 8 >> ss;
        const int n_stripes = (ph + off + SH - 1) / SH, bstride = pw + 8;
        uint16_t *above = (uint16_t *)calloc((size_t)n_stripes * 2 * bstride, 2), *below = (uint16_t *)calloc((size_t)n_stripes * 2 * bstride, 2);
        for (int s = 0; s < n_stripes; s++) {
            const int y0 = s * SH - off > 0 ? s * SH - off : 0, y1 = (s + 1) * SH - off < ph ? (s + 1) * SH - off : ph;
            for (int ab = 0; ab < 2; ab++) { /* 0: above, 1: below */
                uint16_t *dst = (ab ? below : above) + (size_t)s * 2 * bstride + 4;
                const int use_deblock = ab ? (y1 < ph) : (s > 0);
                for (int i = 0; i < 2; i++) {
                    int row;
                    const void *srcp;
                    int sstr;
                    if (use_deblock) { /* save_deblock_boundary_lines: rows y0-2.. / y1.., the last one repeated if short */
                        const int r0 = ab ? y1 : y0 - 2, lines = ph - r0 < 2 ? ph - r0 : 2;
                        row = r0 + (i < lines ? i : 0);
                        srcp = db, sstr = dstr;
                    } else { /* save_cdef_boundary_lines: the outermost CDEF row of the picture, twice */
                        row = ab ? y1 - 1 : y0;
                        srcp = cb, sstr = cs;
                    }
                    for (int x = -4; x < pw + 4; x++)
                        dst[(size_t)i * bstride + x] = (uint16_t)plane_px(srcp, hbd, sstr, x < 0 ? 0 : x >= pw ? pw - 1 : x, row);
                }
            }
        }
        /* units (foreach_rest_unit_in_tile :1366-1411) */
        const int U = rp->restoration_unit_size, ext = U * 3 / 2;
        const int hunits = (pw + (U >> 1)) / U > 1 ? (pw + (U >> 1)) / U : 1;
        uint16_t *res = (uint16_t *)calloc((size_t)ph * pw, 2); /* dst frame */
        int round_0 = 3, round_1 = 11; /* get_conv_params_wiener :104-128 */
        if (bd + 7 - round_0 + 2 > 16) {
            const int e = bd + 7 - round_0 + 2 - 16;
            round_0 += e, round_1 -= e;
        }
        for (int y0u = 0, ui = 0; y0u < ph; ui++) {
            const int rem_h = ph - y0u, uh = rem_h < ext ? rem_h : U;
            int v_start = y0u - off > 0 ? y0u - off : 0, v_end = y0u + uh;
            if (v_end < ph) v_end -= off;
            for (int x0u = 0, uj = 0; x0u < pw; uj++) {
                const int rem_w = pw - x0u, uw = rem_w < ext ? rem_w : U;
                const SvtB200LrUnit *u = &units[plane][ui * hunits + uj];
                if (u->restoration_type == 0) {
                    for (int y = v_start; y < v_end; y++)
                        for (int x = x0u; x < x0u + uw; x++) res[(size_t)y * pw + x] = *wp_at(&wk, x, y);
                } else {
                    for (int i = 0; i < v_end - v_start;) {
                        const int vs = v_start + i;
                        const int first = vs == 0, this_h = SH - (first ? off : 0), last = vs + this_h >= ph;
                        const int copy_above = !first, copy_below = !last;
                        const int stripe = (vs + off) / SH;
                        const int nominal = SH - (stripe == 0 ? off : 0), h = nominal < v_end - vs ? nominal : v_end - vs;
                        /* setup_processing_stripe_boundary: save + overwrite columns x0u-4 .. x0u+uw+4 */
                        uint16_t save[2][3][64 * 6 + 8 + 400];
                        const int lw = uw + 8;
                        if (lw > (int)(sizeof(save[0][0]) / 2)) abort();
                        if (!p->optimized_lr) {
                            if (copy_above)
                                for (int r = -3; r < 0; r++) {
                                    const int br = r + 2 > 0 ? r + 2 : 0;
                                    for (int x = 0; x < lw; x++) {
                                        uint16_t *d = wp_at(&wk, x0u - 4 + x, vs + r);
                                        save[0][r + 3][x] = *d;
                                        *d = above[((size_t)stripe * 2 + br) * bstride + 4 + x0u - 4 + x];
                                    }
                                }
                            if (copy_below)
                                for (int r = 0; r < 3; r++) {
                                    const int br = r < 1 ? r : 1;
                                    for (int x = 0; x < lw; x++) {
                                        uint16_t *d = wp_at(&wk, x0u - 4 + x, vs + h + r);
                                        save[1][r][x] = *d;
                                        *d = below[((size_t)stripe * 2 + br) * bstride + 4 + x0u - 4 + x];
                                    }
                                }
                        } else {
                            if (copy_above)
                                for (int x = 0; x < lw; x++) {
                                    uint16_t *d = wp_at(&wk, x0u - 4 + x, vs - 3);
                                    save[0][0][x] = *d;
                                    *d = *wp_at(&wk, x0u - 4 + x, vs - 2);
                                }
                            if (copy_below)
                                for (int x = 0; x < lw; x++) {
                                    uint16_t *d = wp_at(&wk, x0u - 4 + x, vs + h + 2);
                                    save[1][2][x] = *d;
                                    *d = *wp_at(&wk, x0u - 4 + x, vs + h + 1);
                                }
                        }
                        /* stripe filter, processing units of 64 >> ss_x columns */
                        const int PW = 64 >> ss;
                        for (int j = 0; j < uw; j += PW) {
                            const int w = PW < uw - j ? PW : uw - j;
                            uint16_t tmp[64 * 64];
                            if (u->restoration_type == 1)
                                orc_wiener_convolve_add_src(wp_at(&wk, x0u + j, vs), 1, wk.stride, tmp, 64, u->hfilter, u->vfilter, w, h,
                                                            round_0, round_1, bd);
                            else
                                orc_apply_selfguided_restoration(wp_at(&wk, x0u + j, vs), 1, w, h, wk.stride, u->sgr_ep, u->sgr_xqd, tmp,
                                                                 64, bd);
                            for (int y = 0; y < h; y++)
                                for (int x = 0; x < w; x++) res[(size_t)(vs + y) * pw + x0u + j + x] = tmp[y * 64 + x];
                        }
                        /* restore_processing_stripe_boundary */
                        if (!p->optimized_lr) {
                            if (copy_above)
                                for (int r = -3; r < 0; r++)
                                    for (int x = 0; x < lw; x++) *wp_at(&wk, x0u - 4 + x, vs + r) = save[0][r + 3][x];
                            if (copy_below)
                                for (int r = 0; r < 3; r++) {
                                    if (vs + h + r >= v_end + 3) break;
                                    for (int x = 0; x < lw; x++) *wp_at(&wk, x0u - 4 + x, vs + h + r) = save[1][r][x];
                                }
                        } else {
                            if (copy_above)
                                for (int x = 0; x < lw; x++) *wp_at(&wk, x0u - 4 + x, vs - 3) = save[0][0][x];
                            if (copy_below && vs + h + 2 < v_end + 3)
                                for (int x = 0; x < lw; x++) *wp_at(&wk, x0u - 4 + x, vs + h + 2) = save[1][2][x];
                        }
                        i += h;
                    }
                }
                x0u += uw;
            }
            y0u += uh;
        }
        for (int y = 0; y < ph; y++)
            for (int x = 0; x < pw; x++) {
                if (hbd) ((uint16_t *)ob)[(size_t)y * os + x] = res[(size_t)y * pw + x];
                else ((uint8_t *)ob)[(size_t)y * os + x] = (uint8_t)res[(size_t)y * pw + x];
            }
        free(res);
        free(above);
        free(below);
        free(wk.buf);
    }
}
