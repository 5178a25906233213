/*
 * me_oracle.c — CPU restatement (plain scalar C) of SVT-AV1 v0.8.6 open-loop motion estimation.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (svt-av1_b200/, libsvtav1_b200.so) may
 * include, link or call this file; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs do.  Parity is pinned: every function here is checked bit-for-bit against the reference's own
 * C functions compiled from /root/reference into oracle/_ref/ (tests/test_oracle_vs_ref.py).
 *
 * Each function cites the reference code (relative to /root/reference/Source/Lib/Encoder) it follows.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/svt_av1_b200.h"
#include "oracle.h"

#define ORC_MIN(a, b) ((a) < (b) ? (a) : (b))
#define ORC_MAX(a, b) ((a) > (b) ? (a) : (b))
#define ORC_MAX_SAD_VALUE (128 * 128 * 255) /* Codec/EbMotionEstimation.h:93 */

/* z-order tables, Codec/EbMotionEstimation.h:108-125: raster index -> internal PU index */
static const uint8_t k_tab16[16] = {0, 1, 4, 5, 2, 3, 6, 7, 8, 9, 12, 13, 10, 11, 14, 15};
static const uint8_t k_tab8[64] = {0,  1,  4,  5,  16, 17, 20, 21, 2,  3,  6,  7,  18, 19, 22, 23,
                                   8,  9,  12, 13, 24, 25, 28, 29, 10, 11, 14, 15, 26, 27, 30, 31,
                                   32, 33, 36, 37, 48, 49, 52, 53, 34, 35, 38, 39, 50, 51, 54, 55,
                                   40, 41, 44, 45, 56, 57, 60, 61, 42, 43, 46, 47, 58, 59, 62, 63};

static inline uint32_t absdiff(uint8_t a, uint8_t b) { return a > b ? (uint32_t)(a - b) : (uint32_t)(b - a); }

/* C_DEFAULT/EbComputeSAD_C.c (svt_fast_loop_nxm_sad_kernel / svt_nxm_sad_kernel_helper_c) */
uint32_t orc_nxm_sad(const uint8_t *src, uint32_t src_stride, const uint8_t *ref, uint32_t ref_stride,
                     uint32_t height, uint32_t width) {
    uint32_t sad = 0;
    for (uint32_t y = 0; y < height; y++)
        for (uint32_t x = 0; x < width; x++) sad += absdiff(src[y * src_stride + x], ref[y * ref_stride + x]);
    return sad;
}

/* C_DEFAULT/EbComputeSAD_C.c:57-96 svt_sad_loop_kernel_c: exhaustive search, first minimum in raster
 * order wins (strict <), best starts at 0xffffff; the centre is left untouched if nothing beats it. */
void orc_sad_loop_kernel(const uint8_t *src, uint32_t src_stride, const uint8_t *ref, uint32_t ref_stride,
                         uint32_t block_height, uint32_t block_width, uint64_t *best_sad,
                         int16_t *x_search_center, int16_t *y_search_center, uint32_t src_stride_raw,
                         int16_t search_area_width, int16_t search_area_height) {
    *best_sad = 0xffffff;
    for (int ys = 0; ys < search_area_height; ys++) {
        const uint8_t *row = ref + (size_t)ys * src_stride_raw;
        for (int xs = 0; xs < search_area_width; xs++) {
            uint32_t sad = orc_nxm_sad(src, src_stride, row + xs, ref_stride, block_height, block_width);
            if (sad < *best_sad) {
                *best_sad = sad;
                *x_search_center = (int16_t)xs;
                *y_search_center = (int16_t)ys;
            }
        }
    }
}

/* 8x8 SAD, or with sub_sad the 4 even rows doubled (Codec/EbMotionEstimation.c:66-118,130-160). */
static uint32_t sad8x8(const uint8_t *src, uint32_t ss, const uint8_t *ref, uint32_t rs, int sub_sad) {
    if (sub_sad) return orc_nxm_sad(src, 2 * ss, ref, 2 * rs, 4, 8) << 1;
    return orc_nxm_sad(src, ss, ref, rs, 8, 8);
}

/* Codec/EbMotionEstimation.c:122-186 svt_ext_sad_calculation_8x8_16x16_c (one search point) */
void orc_ext_sad_calculation_8x8_16x16(const uint8_t *src, uint32_t src_stride, const uint8_t *ref,
                                       uint32_t ref_stride, uint32_t *p_best_sad_8x8,
                                       uint32_t *p_best_sad_16x16, uint32_t *p_best_mv8x8,
                                       uint32_t *p_best_mv16x16, uint32_t mv, uint32_t *p_sad16x16,
                                       uint32_t *p_sad8x8, uint8_t sub_sad) {
    uint32_t sum = 0;
    for (int q = 0; q < 4; q++) {
        const uint32_t off_s = (q >> 1) * 8 * src_stride + (q & 1) * 8;
        const uint32_t off_r = (q >> 1) * 8 * ref_stride + (q & 1) * 8;
        p_sad8x8[q] = sad8x8(src + off_s, src_stride, ref + off_r, ref_stride, sub_sad);
        if (p_sad8x8[q] < p_best_sad_8x8[q]) {
            p_best_sad_8x8[q] = p_sad8x8[q];
            p_best_mv8x8[q] = mv;
        }
        sum += p_sad8x8[q];
    }
    if (sum < p_best_sad_16x16[0]) {
        p_best_sad_16x16[0] = sum;
        p_best_mv16x16[0] = mv;
    }
    *p_sad16x16 = sum;
}

/* Codec/EbMotionEstimation.c:191-225 svt_ext_sad_calculation_32x32_64x64_c */
void orc_ext_sad_calculation_32x32_64x64(const uint32_t *p_sad16x16, uint32_t *p_best_sad_32x32,
                                         uint32_t *p_best_sad_64x64, uint32_t *p_best_mv32x32,
                                         uint32_t *p_best_mv64x64, uint32_t mv, uint32_t *p_sad32x32) {
    uint32_t s64 = 0;
    for (int q = 0; q < 4; q++) {
        uint32_t s = p_sad16x16[4 * q] + p_sad16x16[4 * q + 1] + p_sad16x16[4 * q + 2] + p_sad16x16[4 * q + 3];
        p_sad32x32[q] = s;
        if (s < p_best_sad_32x32[q]) {
            p_best_sad_32x32[q] = s;
            p_best_mv32x32[q] = mv;
        }
        s64 += s;
    }
    if (s64 < p_best_sad_64x64[0]) {
        p_best_sad_64x64[0] = s64;
        p_best_mv64x64[0] = mv;
    }
}

/* MV of search point i of an 8-point group: Codec/EbMotionEstimation.c:253-255 */
static inline uint32_t mv_plus(uint32_t mv, int i) {
    int16_t x = (int16_t)((int16_t)(mv & 0xffff) + (int16_t)(i * 4));
    int16_t y = (int16_t)(mv >> 16);
    return ((uint32_t)(uint16_t)y << 16) | (uint16_t)x;
}

/* Codec/EbMotionEstimation.c:230-390 svt_ext_all_sad_calculation_8x8_16x16_c: 8 consecutive x positions,
 * all 64 8x8 and 16 16x16 of a 64x64 SB; per-16x16 order of updates is search index inside block. */
void orc_ext_all_sad_calculation_8x8_16x16(const uint8_t *src, uint32_t src_stride, const uint8_t *ref,
                                           uint32_t ref_stride, uint32_t mv, uint32_t *p_best_sad_8x8,
                                           uint32_t *p_best_sad_16x16, uint32_t *p_best_mv8x8,
                                           uint32_t *p_best_mv16x16, uint32_t p_eight_sad16x16[16][8],
                                           uint32_t p_eight_sad8x8[64][8], uint8_t sub_sad) {
    for (int by = 0; by < 4; by++)
        for (int bx = 0; bx < 4; bx++) {
            const int i16 = k_tab16[4 * by + bx];
            const uint8_t *s = src + 16 * by * src_stride + 16 * bx;
            const uint8_t *r = ref + 16 * by * ref_stride + 16 * bx;
            for (int i = 0; i < 8; i++) {
                uint32_t sum = 0;
                for (int q = 0; q < 4; q++) {
                    const uint32_t os = (q >> 1) * 8 * src_stride + (q & 1) * 8;
                    const uint32_t orf = (q >> 1) * 8 * ref_stride + (q & 1) * 8;
                    uint32_t v = sad8x8(s + os, src_stride, r + orf + i, ref_stride, sub_sad);
                    p_eight_sad8x8[4 * i16 + q][i] = v;
                    if (v < p_best_sad_8x8[4 * i16 + q]) {
                        p_best_sad_8x8[4 * i16 + q] = v;
                        p_best_mv8x8[4 * i16 + q] = mv_plus(mv, i);
                    }
                    sum += v;
                }
                p_eight_sad16x16[i16][i] = sum;
                if (sum < p_best_sad_16x16[i16]) {
                    p_best_sad_16x16[i16] = sum;
                    p_best_mv16x16[i16] = mv_plus(mv, i);
                }
            }
        }
}

/* Codec/EbMotionEstimation.c:396-455 svt_ext_eight_sad_calculation_32x32_64x64_c */
void orc_ext_eight_sad_calculation_32x32_64x64(uint32_t p_sad16x16[16][8], uint32_t *p_best_sad_32x32,
                                               uint32_t *p_best_sad_64x64, uint32_t *p_best_mv32x32,
                                               uint32_t *p_best_mv64x64, uint32_t mv,
                                               uint32_t p_sad32x32[4][8]) {
    for (int i = 0; i < 8; i++) {
        uint32_t s64 = 0;
        for (int q = 0; q < 4; q++) {
            uint32_t s = p_sad16x16[4 * q][i] + p_sad16x16[4 * q + 1][i] + p_sad16x16[4 * q + 2][i] +
                         p_sad16x16[4 * q + 3][i];
            p_sad32x32[q][i] = s;
            if (s < p_best_sad_32x32[q]) {
                p_best_sad_32x32[q] = s;
                p_best_mv32x32[q] = mv_plus(mv, i);
            }
            s64 += s;
        }
        if (s64 < p_best_sad_64x64[0]) {
            p_best_sad_64x64[0] = s64;
            p_best_mv64x64[0] = mv_plus(mv, i);
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * SB-level driver
 * ---------------------------------------------------------------------------------------------- */

/* The reference's window clamp, statement for statement (e.g. Codec/EbMotionEstimation.c:927-975):
 * NB the left/top clamp moves the origin FIRST and then re-tests the (already corrected) origin, so the
 * size is never reduced on that side; only the right/bottom clamp shrinks it. */
static void clamp_window(int origin, int pad, int pic_size, int *ao, int *size) {
    if (origin + *ao < -pad) *ao = -pad - origin;
    if (origin + *ao < -pad) *size = *size - (-pad - (origin + *ao)); /* never true; kept for fidelity */
    if (origin + *ao > pic_size - 1) *ao = *ao - ((origin + *ao) - (pic_size - 1));
    if (origin + *ao + *size > pic_size) *size = ORC_MAX(1, *size - ((origin + *ao + *size) - pic_size));
}

static int scaled_dist(int dist) { /* ((dist*5)/8) + round_up, EbMotionEstimation.c:1928-1931,2262-2265 */
    int round_up = (dist % 8) == 0 ? 0 : 1;
    return ((dist * 5) / 8) + round_up;
}

typedef struct {
    uint64_t sad;
    int16_t x, y;
} LevelRes;

/* Codec/EbMotionEstimation.c:852-1023 hme_level_0 (TWO_DECIMATION_HME) */
static LevelRes hme_l0(const SvtB200MeParams *p, const uint8_t *src16, const uint8_t *ref16, int ox, int oy,
                       int sbw, int sbh, int rx, int ry, int mult) {
    const SvtB200Plane *pl = &p->sixteenth;
    int saw = ORC_MIN((int16_t)(((p->hme_level0_search_area_in_width_array[rx] * mult) / 100 + 15) & ~0x0F),
                      (int16_t)((p->hme_level0_max_search_area_in_width_array[rx] + 15) & ~0x0F));
    int sah = ORC_MIN((int16_t)((p->hme_level0_search_area_in_height_array[ry] * mult) / 100),
                      (int16_t)p->hme_level0_max_search_area_in_height_array[ry]);
    int xdist = 0, ydist = 0;
    for (int i = rx; i-- > 0;)
        xdist += ORC_MIN((int16_t)((p->hme_level0_search_area_in_width_array[i] * mult) / 100),
                         (int16_t)p->hme_level0_max_search_area_in_width_array[i]);
    for (int i = ry; i-- > 0;)
        ydist += ORC_MIN((int16_t)((p->hme_level0_search_area_in_height_array[i] * mult) / 100),
                         (int16_t)p->hme_level0_max_search_area_in_height_array[i]);
    int xo = -(int16_t)(ORC_MIN((p->hme_level0_total_search_area_width * mult) / 100,
                                p->hme_level0_max_total_search_area_width) >> 1) + xdist;
    int yo = -(int16_t)(ORC_MIN((p->hme_level0_total_search_area_height * mult) / 100,
                                p->hme_level0_max_total_search_area_height) >> 1) + ydist;
    clamp_window(ox, pl->origin_x - 1, pl->width, &xo, &saw);
    saw = saw < 16 ? saw : saw & ~0x0F;
    clamp_window(oy, pl->origin_y - 1, pl->height, &yo, &sah);
    const int sub = p->hme_search_method != 0;
    const uint8_t *s = src16 + (pl->origin_y + oy) * pl->stride + pl->origin_x + ox;
    const uint8_t *r = ref16 + (pl->origin_y + oy + yo) * pl->stride + pl->origin_x + ox + xo;
    LevelRes res = {0, 0, 0};
    /* the process kernel pre-gathers every other row of the SB into sixteenth_sb_buffer
     * (EbMotionEstimationProcess.c:893-908); reading the plane with a doubled stride is the same */
    orc_sad_loop_kernel(s, sub ? pl->stride * 2 : pl->stride, r, sub ? pl->stride * 2 : pl->stride,
                        sub ? sbh >> 1 : sbh, sbw, &res.sad, &res.x, &res.y, pl->stride, (int16_t)saw,
                        (int16_t)sah);
    if (sub) res.sad *= 2;
    res.x = (int16_t)((res.x + xo) * 4);
    res.y = (int16_t)((res.y + yo) * 4);
    return res;
}

/* Codec/EbMotionEstimation.c:1028-1175 hme_level_1 and :1177-1318 hme_level_2 (TWO_DECIMATION_HME:
 * the areas are not rescaled by the distance factor).  `scale` is 2 for level 1, 1 for level 2. */
static LevelRes hme_l12(const SvtB200MeParams *p, const SvtB200Plane *pl, int pad_x, int pad_y,
                        const uint8_t *srcp, const uint8_t *refp, int ox, int oy, int sbw, int sbh, int aw,
                        int ah, int cx, int cy, int scale) {
    int saw = (int16_t)((aw + 7) & ~0x07);
    int sah = ah;
    int xo = -(saw >> 1) + cx;
    int yo = -(sah >> 1) + cy;
    clamp_window(ox, pad_x, pl->width, &xo, &saw);
    saw = saw < 8 ? saw : saw & ~0x07;
    clamp_window(oy, pad_y, pl->height, &yo, &sah);
    const int sub = p->hme_search_method != 0;
    const uint8_t *s = srcp + (pl->origin_y + oy) * pl->stride + pl->origin_x + ox;
    const uint8_t *r = refp + (pl->origin_y + oy + yo) * pl->stride + pl->origin_x + ox + xo;
    LevelRes res = {0, 0, 0};
    orc_sad_loop_kernel(s, sub ? pl->stride * 2 : pl->stride, r, sub ? pl->stride * 2 : pl->stride,
                        sub ? sbh >> 1 : sbh, sbw, &res.sad, &res.x, &res.y, pl->stride, (int16_t)saw,
                        (int16_t)sah);
    if (sub) res.sad *= 2;
    res.x = (int16_t)((res.x + xo) * scale);
    res.y = (int16_t)((res.y + yo) * scale);
    return res;
}

/* Stable selection-by-swap used by hme_prune_ref_and_adjust_sr / me_prune_ref only to find the minimum
 * hme_sad over all 2x4 slots (Codec/EbMotionEstimation.c:2176-2186, 2784-2794): best = min. */
static uint64_t min_hme_sad(const SvtB200HmeResult h[2][4]) {
    uint64_t best = h[0][0].hme_sad;
    for (int l = 0; l < 2; l++)
        for (int r = 0; r < 4; r++)
            if (h[l][r].hme_sad < best) best = h[l][r].hme_sad;
    return best;
}

/* motion_estimate_sb (Codec/EbMotionEstimation.c:2912-3041) for one 64x64 SB. */
void orc_me_sb(const SvtB200MeParams *p, const SvtB200MePlanes *src,
               const SvtB200MePlanes refs[2][4], int sb_x, int sb_y, uint32_t *best_sad /*[2][4][85]*/,
               uint32_t *best_mv /*[2][4][85]*/, SvtB200HmeResult *hme_out /*[2][4]*/,
               int16_t *me_mv /*[85*7*2]*/, uint8_t *me_cand /*[85*23]*/, uint8_t *total_cand /*[85]*/,
               uint32_t *rc_me_distortion) {
    const int ox = sb_x * 64, oy = sb_y * 64;
    const int pic_w = p->full.width, pic_h = p->full.height;
    const int sbw = ORC_MIN(pic_w - ox, 64), sbh = ORC_MIN(pic_h - oy, 64);
    SvtB200HmeResult hme[2][4];
    uint32_t divisor[2][4];
    uint32_t(*bsad)[4][85] = (uint32_t(*)[4][85])best_sad;
    uint32_t(*bmv)[4][85] = (uint32_t(*)[4][85])best_mv;

    /* :2928-2943 init */
    for (int l = 0; l < 2; l++)
        for (int r = 0; r < 4; r++) {
            hme[l][r].sc_x = hme[l][r].sc_y = 0;
            hme[l][r].do_ref = 1;
            hme[l][r].hme_sad = 0xFFFFFFFFull;
            divisor[l][r] = 1;
            for (int i = 0; i < 85; i++) {
                bmv[l][r][i] = 0;
                bsad[l][r][i] = 0; /* the reference leaves stale data; we define 0 for slots never searched */
            }
        }

    /* hme_sb: :2744-2778, levels :2204-2570, final centre :2575-2740 */
    const int nrw = p->number_hme_search_region_in_width, nrh = p->number_hme_search_region_in_height;
    uint64_t hme_mv_sad = 0; /* NB: carried across refs exactly like the reference's local (:2589) */
    int16_t xc = 0, yc = 0; /* xHmeSearchCenter/yHmeSearchCenter are also function-scope locals */
    for (int l = 0; l < p->num_lists; l++)
        for (int r = 0; r < p->num_refs[l]; r++) {
            int16_t scx = 0, scy = 0;
            if (p->temporal_layer_index > 0 || l == 0) {
                if (p->enable_hme_flag) {
                    LevelRes l0[2][2], l1[2][2], l2[2][2];
                    const int mult = scaled_dist(p->ref_dist[l][r]) * 100;
                    for (int ry = 0; ry < nrh; ry++)
                        for (int rx = 0; rx < nrw; rx++) {
                            LevelRes z = {0, 0, 0};
                            l0[rx][ry] = l1[rx][ry] = l2[rx][ry] = z;
                            if (p->enable_hme_level0_flag)
                                l0[rx][ry] = hme_l0(p, src->sixteenth, refs[l][r].sixteenth, ox >> 2, oy >> 2,
                                                    sbw >> 2, sbh >> 2, rx, ry, mult);
                            l1[rx][ry] = l0[rx][ry]; /* levels that do not run keep the previous centre */
                            if (p->enable_hme_level1_flag)
                                l1[rx][ry] = hme_l12(p, &p->quarter, p->quarter.origin_x - 1,
                                                     p->quarter.origin_y - 1, src->quarter, refs[l][r].quarter,
                                                     ox >> 1, oy >> 1, sbw >> 1, sbh >> 1,
                                                     p->hme_level1_search_area_in_width_array[rx],
                                                     p->hme_level1_search_area_in_height_array[ry],
                                                     l0[rx][ry].x >> 1, l0[rx][ry].y >> 1, 2);
                            l2[rx][ry] = l1[rx][ry];
                            if (p->enable_hme_level2_flag)
                                l2[rx][ry] = hme_l12(p, &p->full, 63, 63, src->full, refs[l][r].full, ox, oy,
                                                     sbw, sbh, p->hme_level2_search_area_in_width_array[rx],
                                                     p->hme_level2_search_area_in_height_array[ry],
                                                     l1[rx][ry].x, l1[rx][ry].y, 1);
                        }
                    /* set_final_seach_centre_sb: last enabled level, first strict minimum over regions
                     * visited in (ry outer, rx inner) order starting from region (0,0) */
                    LevelRes(*lv)[2] = p->enable_hme_level2_flag ? l2 : p->enable_hme_level1_flag ? l1 : l0;
                    if (p->enable_hme_level0_flag || p->enable_hme_level1_flag || p->enable_hme_level2_flag) {
                        xc = lv[0][0].x;
                        yc = lv[0][0].y;
                        hme_mv_sad = lv[0][0].sad;
                        for (int ry = 0; ry < nrh; ry++)
                            for (int rx = (ry == 0 ? 1 : 0); rx < nrw; rx++)
                                if (lv[rx][ry].sad < hme_mv_sad) {
                                    xc = lv[rx][ry].x;
                                    yc = lv[rx][ry].y;
                                    hme_mv_sad = lv[rx][ry].sad;
                                }
                    }
                    scx = xc;
                    scy = yc;
                }
            }
            hme[l][r].sc_x = scx;
            hme[l][r].sc_y = scy;
            hme[l][r].hme_sad = hme_mv_sad;
            hme[l][r].do_ref = 1;
        }

    /* :2948-2952 hme_prune_ref_and_adjust_sr (:2779-2823) */
    const int prune_ref = p->enable_hme_flag && p->enable_hme_level2_flag; /* me_type != ME_MCTF */
    if (prune_ref && (p->enable_me_sr_adjustment || p->enable_me_hme_ref_pruning)) {
        const uint64_t best = min_hme_sad(hme);
        const uint16_t th = (uint16_t)p->prune_ref_if_hme_sad_dev_bigger_than_th;
        for (int l = 0; l < 2; l++)
            for (int r = 0; r < 4; r++) {
                if (p->enable_me_hme_ref_pruning && th != 0xFFFF && (hme[l][r].hme_sad - best) * 100 > th * best)
                    hme[l][r].do_ref = 0;
                if (p->enable_me_sr_adjustment) {
                    if (abs(hme[l][r].sc_x) <= p->reduce_me_sr_based_on_mv_length_th &&
                        abs(hme[l][r].sc_y) <= p->reduce_me_sr_based_on_mv_length_th &&
                        hme[l][r].hme_sad < (uint64_t)p->stationary_hme_sad_abs_th)
                        divisor[l][r] = p->stationary_me_sr_divisor;
                    else if (hme[l][r].hme_sad < (uint64_t)p->reduce_me_sr_based_on_hme_sad_abs_th)
                        divisor[l][r] = p->me_sr_divisor_for_low_hme_sad;
                }
            }
    }

    /* integer_search_sb :1868-2139 (unrestricted MVs, single tile) */
    const int sub_me = p->me_search_method != 0;
    const SvtB200Plane *fp = &p->full;
    for (int l = 0; l < p->num_lists; l++)
        for (int r = 0; r < p->num_refs[l]; r++) {
            if (!hme[l][r].do_ref) continue;
            int16_t xsc = hme[l][r].sc_x, ysc = hme[l][r].sc_y;
            const int dist = scaled_dist(p->ref_dist[l][r]);
            int saw = (int16_t)ORC_MIN(p->search_area_width * dist, p->max_me_search_width);
            int sah = (int16_t)ORC_MIN(p->search_area_height * dist, p->max_me_search_height);
            saw = (int16_t)(((saw / (int)divisor[l][r]) + 7) & ~0x07);
            sah = (int16_t)ORC_MAX(1, sah / (int)divisor[l][r]);
            const uint8_t *srcb = src->full + (fp->origin_y + oy) * fp->stride + fp->origin_x + ox;
            const uint8_t *refb = refs[l][r].full + (fp->origin_y + oy) * fp->stride + fp->origin_x + ox;
            if ((xsc != 0 || ysc != 0) && p->is_used_as_reference_flag) {
                /* check_00_center :1348-1421 */
                uint32_t zero_sad = orc_nxm_sad(srcb, fp->stride << 1, refb, fp->stride << 1, sbh >> 1, sbw) << 1;
                int cx = xsc, cy = ysc;
                if (ox + cx < -63) cx = -63 - ox;
                if (ox + cx > fp->width - 1) cx = cx - ((ox + cx) - (fp->width - 1));
                if (oy + cy < -63) cy = -63 - oy;
                if (oy + cy > fp->height - 1) cy = cy - ((oy + cy) - (fp->height - 1));
                uint32_t hme_sad = orc_nxm_sad(srcb, fp->stride << 1, refb + cy * fp->stride + cx,
                                               fp->stride << 1, sbh >> 1, sbw) << 1;
                if (zero_sad <= hme_sad) cx = cy = 0;
                xsc = (int16_t)cx;
                ysc = (int16_t)cy;
            }
            int xo = xsc - (saw >> 1), yo = ysc - (sah >> 1);
            clamp_window(ox, 63, pic_w, &xo, &saw);
            saw = saw < 8 ? saw : saw & ~0x07;
            clamp_window(oy, 63, pic_h, &yo, &sah);
            for (int i = 0; i < 85; i++) bsad[l][r][i] = ORC_MAX_SAD_VALUE;
            /* open_loop_me_fullpel_search_sblock :814-850: groups of 8 then single-point tail */
            uint32_t e16[16][8], e8[64][8], e32[4][8], s16[16], s8[64], s32[4];
            const int w8 = saw & ~7;
            for (int ys = 0; ys < sah; ys++) {
                const uint8_t *rrow = refb + (yo + ys) * fp->stride + xo;
                for (int xs = 0; xs < w8; xs += 8) {
                    uint32_t mv = ((uint32_t)(uint16_t)(yo + ys) << 18) | (uint16_t)((uint16_t)(xo + xs) << 2);
                    orc_ext_all_sad_calculation_8x8_16x16(srcb, fp->stride, rrow + xs, fp->stride, mv,
                                                          &bsad[l][r][21], &bsad[l][r][5], &bmv[l][r][21],
                                                          &bmv[l][r][5], e16, e8, (uint8_t)sub_me);
                    orc_ext_eight_sad_calculation_32x32_64x64(e16, &bsad[l][r][1], &bsad[l][r][0],
                                                              &bmv[l][r][1], &bmv[l][r][0], mv, e32);
                }
                for (int xs = w8; xs < saw; xs++) {
                    uint32_t mv = ((uint32_t)(uint16_t)(yo + ys) << 18) | (uint16_t)((uint16_t)(xo + xs) << 2);
                    for (int b = 0; b < 16; b++) { /* :555-795, blocks visited in internal-index order */
                        int raster = 0;
                        for (int k = 0; k < 16; k++)
                            if (k_tab16[k] == b) raster = k;
                        const int by = raster >> 2, bx = raster & 3;
                        orc_ext_sad_calculation_8x8_16x16(srcb + 16 * by * fp->stride + 16 * bx, fp->stride,
                                                          rrow + xs + 16 * by * fp->stride + 16 * bx, fp->stride,
                                                          &bsad[l][r][21 + 4 * b], &bsad[l][r][5 + b],
                                                          &bmv[l][r][21 + 4 * b], &bmv[l][r][5 + b], mv, &s16[b],
                                                          &s8[4 * b], (uint8_t)sub_me);
                    }
                    orc_ext_sad_calculation_32x32_64x64(s16, &bsad[l][r][1], &bsad[l][r][0], &bmv[l][r][1],
                                                        &bmv[l][r][0], mv, s32);
                }
            }
        }

    /* me_prune_ref :2145-2199 */
    if (prune_ref && p->enable_me_hme_ref_pruning) {
        for (int l = 0; l < p->num_lists; l++)
            for (int r = 0; r < p->num_refs[l]; r++) {
                if (!hme[l][r].do_ref) {
                    hme[l][r].hme_sad = (uint64_t)ORC_MAX_SAD_VALUE * 64;
                    continue;
                }
                uint64_t s = 0;
                for (int i = 0; i < 64; i++) s += bsad[l][r][21 + i];
                hme[l][r].hme_sad = s;
            }
        const uint64_t best = min_hme_sad(hme);
        const uint16_t th = (uint16_t)p->prune_ref_if_me_sad_dev_bigger_than_th;
        for (int l = 0; l < 2; l++)
            for (int r = 0; r < 4; r++)
                if (th != 0xFFFF && (hme[l][r].hme_sad - best) * 100 > th * best) hme[l][r].do_ref = 0;
    }
    memcpy(hme_out, hme, sizeof(hme));

    /* candidates + MeSbResults :2964-3021, construct_me_candidate_array :2825-2905 */
    memset(me_mv, 0, sizeof(int16_t) * 85 * 7 * 2);
    memset(me_cand, 0, 85 * 23);
    memset(total_cand, 0, 85);
    uint32_t first_cand_dist[85];
    memset(first_cand_dist, 0, sizeof(first_cand_dist));
    for (int pu = 0; pu < p->max_number_of_pus_per_sb; pu++) {
        const int n_idx = pu > 20 ? k_tab8[pu - 21] + 21 : pu > 4 ? k_tab16[pu - 5] + 5 : pu;
        int n = 0;
        uint8_t *cand = me_cand + pu * 23;
        for (int l = 0; l < p->num_lists; l++)
            for (int r = 0; r < p->num_refs[l]; r++) {
                if (!hme[l][r].do_ref) continue;
                if (n == 0) first_cand_dist[pu] = bsad[l][r][n_idx];
                /* direction=l, ref_index[l]=r, ref0_list = (l==0 ? 0 : 24)&1, ref1_list = (l==1 ? 1 : 24)&1 */
                if (n < 23) cand[n] = (uint8_t)(l | (l == 0 ? (r << 2) : (r << 4)) | (l == 1 ? 0x80 : 0));
                n++;
            }
        if (p->num_lists > 1) {
            for (int a = 0; a < p->num_refs[0]; a++)
                for (int b = 0; b < p->num_refs[1]; b++)
                    if (hme[0][a].do_ref && hme[1][b].do_ref) {
                        if (n < 23) cand[n] = (uint8_t)(2 | (a << 2) | (b << 4) | 0x80);
                        n++;
                    }
            for (int a = 1; a < p->num_refs[0]; a++)
                if (hme[0][0].do_ref && hme[0][a].do_ref) {
                    if (n < 23) cand[n] = (uint8_t)(2 | (0 << 2) | (a << 4));
                    n++;
                }
            if (p->num_refs[1] == 3 && hme[1][0].do_ref && hme[1][2].do_ref) {
                if (n < 23) cand[n] = (uint8_t)(2 | (0 << 2) | (2 << 4) | 0x40 | 0x80);
                n++;
            }
        }
        total_cand[pu] = (uint8_t)ORC_MIN(n, 23);
        for (int l = 0; l < p->num_lists; l++)
            for (int r = 0; r < p->num_refs[l]; r++) {
                const int slot = pu * 7 + (l ? 4 : 0) + r;
                me_mv[2 * slot] = (int16_t)(bmv[l][r][n_idx] & 0xffff);
                me_mv[2 * slot + 1] = (int16_t)(bmv[l][r][n_idx] >> 16);
            }
    }
    uint32_t rc = 0;
    if (p->rc_dist_from_8x8)
        for (int i = 0; i < 64; i++) rc += first_cand_dist[21 + i];
    else
        for (int i = 0; i < 16; i++) rc += first_cand_dist[5 + i];
    *rc_me_distortion = rc;
}

void orc_me_picture(const SvtB200MeParams *p, const SvtB200MePlanes *src, const SvtB200MePlanes refs[2][4],
                    uint32_t *best_sad, uint32_t *best_mv, SvtB200HmeResult *hme, int16_t *me_mv,
                    uint8_t *me_cand, uint8_t *total_cand, uint32_t *rc_me_distortion) {
    const int sbs_x = (p->full.width + 63) / 64, sbs_y = (p->full.height + 63) / 64;
    for (int sy = 0; sy < sbs_y; sy++)
        for (int sx = 0; sx < sbs_x; sx++) {
            const int sb = sy * sbs_x + sx;
            orc_me_sb(p, src, refs, sx, sy, best_sad + (size_t)sb * 2 * 4 * 85, best_mv + (size_t)sb * 2 * 4 * 85,
                      hme + (size_t)sb * 8, me_mv + (size_t)sb * 85 * 7 * 2, me_cand + (size_t)sb * 85 * 23,
                      total_cand + (size_t)sb * 85, rc_me_distortion + sb);
        }
}

/* ---------------------------------------------------------------------------------------------------------------------
 * The 1/4 and 1/16 luma planes HME searches (picture-analysis decimation).
 * Follows Source/Lib/Encoder/Codec/EbPictureAnalysisProcess.c: downsample_2d :223-256 (2x2 average, rounding),
 * decimation_2d :193-216 (top-left sample), downsample_filtering_input_picture[_ime] :3606-3720 (quarter from full,
 * sixteenth from quarter) / downsample_decimation_input_picture[_ime] :3312-3360 (both from full), followed by
 * generate_padding (Common/Codec/EbMcp.c:112-164: rows first, then whole padded rows up and down).
 * planes->full holds the picture (its own padding is not touched); the other two planes are written in full.
 * ------------------------------------------------------------------------------------------------------------------- */
static void orc_pad_plane(uint8_t *buf, const SvtB200Plane *g) {
    for (int y = 0; y < g->height; y++) {
        uint8_t *row = buf + (size_t)(g->origin_y + y) * g->stride + g->origin_x;
        memset(row - g->origin_x, row[0], (size_t)g->origin_x);
        memset(row + g->width, row[g->width - 1], (size_t)g->origin_x);
    }
    for (int k = 1; k <= g->origin_y; k++) {
        memcpy(buf + (size_t)(g->origin_y - k) * g->stride, buf + (size_t)g->origin_y * g->stride, (size_t)g->stride);
        memcpy(buf + (size_t)(g->origin_y + g->height - 1 + k) * g->stride, buf + (size_t)(g->origin_y + g->height - 1) * g->stride,
               (size_t)g->stride);
    }
}
static void orc_shrink(const uint8_t *in, int in_stride, int in_w, int in_h, uint8_t *out, int out_stride, int step, int filtered) {
    const int half = step >> 1;
    if (filtered) {
        for (int y = half, oy = 0; y < in_h; y += step, oy++)
            for (int x = half, ox = 0; x < in_w; x += step, ox++)
                out[oy * out_stride + ox] = (uint8_t)((in[(y - 1) * in_stride + x - 1] + in[(y - 1) * in_stride + x] +
                                                       in[y * in_stride + x - 1] + in[y * in_stride + x] + 2) >> 2);
    } else {
        for (int y = 0; y < in_h; y += step)
            for (int x = 0; x < in_w; x += step) out[(y / step) * out_stride + x / step] = in[y * in_stride + x];
    }
}
ORC_API void orc_me_downsample(const SvtB200Plane *full, const SvtB200Plane *quarter, const SvtB200Plane *sixteenth,
                               const SvtB200MePlanes *planes, int filtered) {
    const uint8_t *f = (const uint8_t *)planes->full + (size_t)full->origin_y * full->stride + full->origin_x;
    uint8_t *q = (uint8_t *)planes->quarter + (size_t)quarter->origin_y * quarter->stride + quarter->origin_x;
    uint8_t *s = (uint8_t *)planes->sixteenth + (size_t)sixteenth->origin_y * sixteenth->stride + sixteenth->origin_x;
    orc_shrink(f, full->stride, full->width, full->height, q, quarter->stride, 2, filtered);
    orc_pad_plane((uint8_t *)planes->quarter, quarter);
    if (filtered)
        orc_shrink(q, quarter->stride, quarter->width, quarter->height, s, sixteenth->stride, 2, 1);
    else
        orc_shrink(f, full->stride, full->width, full->height, s, sixteenth->stride, 4, 0);
    orc_pad_plane((uint8_t *)planes->sixteenth, sixteenth);
}
