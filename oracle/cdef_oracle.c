/*
 * cdef_oracle.c — CPU restatement (plain scalar C) of SVT-AV1 v0.8.6 CDEF: direction search, the
 * directional filter, the per-64x64 strength search and the frame apply.
 *
 * TEST INFRASTRUCTURE ONLY (see me_oracle.c).  Pinned bit-for-bit against the reference's own C functions
 * (svt_cdef_find_dir_c, svt_cdef_filter_block_c, svt_cdef_filter_fb, compute_cdef_dist*, cdef_seg_search,
 * svt_av1_cdef_frame) compiled into oracle/_ref — tests/test_oracle_cdef.py.
 * Reference paths are relative to /root/reference/Source/Lib.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/svt_av1_b200.h"
#include "oracle.h"

#define VERY_LARGE 16384 /* CDEF_VERY_LARGE, Common/Codec/EbCdef.h:37 */

static int msb(uint32_t n) { /* get_msb: index of the highest set bit */
    int r = 0;
    while (n >>= 1) r++;
    return r;
}

/* Common/Codec/EbCdef.c:86-93 constrain() */
static int constrain(int diff, int threshold, int damping) {
    if (!threshold) return 0;
    int shift = damping - msb((uint32_t)threshold);
    if (shift < 0) shift = 0;
    int mag = abs(diff), lim = threshold - (mag >> shift);
    if (lim < 0) lim = 0;
    if (mag > lim) mag = lim;
    return diff < 0 ? -mag : mag;
}

/* Common/Codec/EbCdef.c:113-117 adjust_strength() */
static int adjust_strength(int strength, int var) {
    const int i = (var >> 6) ? (msb((uint32_t)(var >> 6)) < 12 ? msb((uint32_t)(var >> 6)) : 12) : 0;
    return var ? (strength * (4 + i) + 8) >> 4 : 0;
}

/* Common/Codec/EbCdef.c:132-197 svt_cdef_find_dir_c */
int32_t orc_cdef_find_dir(const uint16_t *img, int32_t stride, int32_t *var, int32_t coeff_shift) {
    static const int div_table[9] = {0, 840, 420, 280, 210, 168, 140, 120, 105};
    int cost[8] = {0}, partial[8][15];
    memset(partial, 0, sizeof(partial));
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            const int x = (img[i * stride + j] >> coeff_shift) - 128;
            partial[0][i + j] += x;
            partial[1][i + j / 2] += x;
            partial[2][i] += x;
            partial[3][3 + i - j / 2] += x;
            partial[4][7 + i - j] += x;
            partial[5][3 - i / 2 + j] += x;
            partial[6][j] += x;
            partial[7][i / 2 + j] += x;
        }
    for (int i = 0; i < 8; i++) {
        cost[2] += partial[2][i] * partial[2][i];
        cost[6] += partial[6][i] * partial[6][i];
    }
    cost[2] *= div_table[8];
    cost[6] *= div_table[8];
    for (int i = 0; i < 7; i++) {
        cost[0] += (partial[0][i] * partial[0][i] + partial[0][14 - i] * partial[0][14 - i]) * div_table[i + 1];
        cost[4] += (partial[4][i] * partial[4][i] + partial[4][14 - i] * partial[4][14 - i]) * div_table[i + 1];
    }
    cost[0] += partial[0][7] * partial[0][7] * div_table[8];
    cost[4] += partial[4][7] * partial[4][7] * div_table[8];
    for (int d = 1; d < 8; d += 2) {
        for (int j = 0; j < 5; j++) cost[d] += partial[d][3 + j] * partial[d][3 + j];
        cost[d] *= div_table[8];
        for (int j = 0; j < 3; j++)
            cost[d] += (partial[d][j] * partial[d][j] + partial[d][10 - j] * partial[d][10 - j]) * div_table[2 * j + 2];
    }
    int best_cost = 0, best_dir = 0;
    for (int d = 0; d < 8; d++)
        if (cost[d] > best_cost) {
            best_cost = cost[d];
            best_dir = d;
        }
    *var = (best_cost - cost[(best_dir + 4) & 7]) >> 10;
    return best_dir;
}

/* direction offsets (dy, dx) for tap k = 0,1 — eb_cdef_directions, Common/Codec/EbCdef.c:96-104 */
static const int8_t k_dir[8][2][2] = {{{-1, 1}, {-2, 2}}, {{0, 1}, {-1, 2}}, {{0, 1}, {0, 2}}, {{0, 1}, {1, 2}},
                                      {{1, 1}, {2, 2}},   {{1, 0}, {2, 1}},  {{1, 0}, {2, 0}}, {{1, 0}, {2, -1}}};

/* One filtered sample. `in` points at the sample inside a tile of row stride `s`; samples equal to VERY_LARGE
 * are "unavailable". Common/Codec/EbCdef.c:202-257 (body of svt_cdef_filter_block_c). */
static int cdef_sample(const uint16_t *in, int s, int pri_strength, int sec_strength, int dir, int pri_damping,
                       int sec_damping, int coeff_shift) {
    static const int pri_taps[2][2] = {{4, 2}, {3, 3}};
    static const int sec_taps[2][2] = {{2, 1}, {2, 1}};
    const int *pt = pri_taps[(pri_strength >> coeff_shift) & 1];
    const int *st = sec_taps[(pri_strength >> coeff_shift) & 1];
    int16_t sum = 0;
    const int16_t x = (int16_t)in[0];
    int mx = x, mn = x;
    for (int k = 0; k < 2; k++) {
        const int po = k_dir[dir][k][0] * s + k_dir[dir][k][1];
        const int16_t p[2] = {(int16_t)in[po], (int16_t)in[-po]};
        for (int t = 0; t < 2; t++) {
            sum = (int16_t)(sum + (int16_t)(pt[k] * constrain(p[t] - x, pri_strength, pri_damping)));
            if (p[t] != VERY_LARGE && p[t] > mx) mx = p[t];
            if (p[t] < mn) mn = p[t];
        }
        const int o2 = k_dir[(dir + 2) & 7][k][0] * s + k_dir[(dir + 2) & 7][k][1];
        const int o6 = k_dir[(dir + 6) & 7][k][0] * s + k_dir[(dir + 6) & 7][k][1];
        const int16_t q[4] = {(int16_t)in[o2], (int16_t)in[-o2], (int16_t)in[o6], (int16_t)in[-o6]};
        for (int t = 0; t < 4; t++) {
            if (q[t] != VERY_LARGE && q[t] > mx) mx = q[t];
            if (q[t] < mn) mn = q[t];
        }
        for (int t = 0; t < 4; t++)
            sum = (int16_t)(sum + (int16_t)(st[k] * constrain(q[t] - x, sec_strength, sec_damping)));
    }
    int y = x + ((8 + sum - (sum < 0)) >> 4);
    y = y < mn ? mn : (y > mx ? mx : y);
    return (int16_t)y;
}

/* svt_cdef_filter_block_c. bsize: BLOCK_4X4=0, BLOCK_4X8=1, BLOCK_8X4=2, BLOCK_8X8=3 (EbDefinitions.h BlockSize) */
void orc_cdef_filter_block(uint8_t *dst8, uint16_t *dst16, int32_t dstride, const uint16_t *in,
                           int32_t pri_strength, int32_t sec_strength, int32_t dir, int32_t pri_damping,
                           int32_t sec_damping, int32_t bsize, int32_t coeff_shift) {
    const int rows = 4 << (bsize == 3 || bsize == 1), cols = 4 << (bsize == 3 || bsize == 2);
    for (int i = 0; i < rows; i++)
        for (int j = 0; j < cols; j++) {
            int y = cdef_sample(in + i * 144 + j, 144, pri_strength, sec_strength, dir, pri_damping, sec_damping,
                                coeff_shift);
            if (dst8)
                dst8[i * dstride + j] = (uint8_t)y;
            else
                dst16[i * dstride + j] = (uint16_t)y;
        }
}

/* ---- picture-level helpers ------------------------------------------------------------------------ */
typedef struct {
    const void *p;
    int stride, w, h, hbd;
} PlaneRO;

static PlaneRO plane_of(const SvtB200Frame *f, int pli, int mi_rows, int mi_cols) {
    PlaneRO r;
    r.p = pli == 0 ? f->y : pli == 1 ? f->cb : f->cr;
    r.stride = pli == 0 ? f->stride_y : f->stride_c;
    r.w = (mi_cols * 4) >> (pli != 0);
    r.h = (mi_rows * 4) >> (pli != 0);
    r.hbd = f->bit_depth > 8;
    return r;
}
static inline int px(const PlaneRO *p, int y, int x) {
    return p->hbd ? ((const uint16_t *)p->p)[(size_t)y * p->stride + x] : ((const uint8_t *)p->p)[(size_t)y * p->stride + x];
}

/* Build the filter-block input tile the way cdef_seg_search does (EbCdefProcess.c:205-226): everything
 * VERY_LARGE, then the block plus the borders that lie inside the frame. tile: [(bh+4)][tile stride ts], origin at
 * (2,2). (The reference uses 3-row/8-column borders; the filter reaches at most 2.) */
static void load_tile(const PlaneRO *pl, int y0, int x0, int bh, int bw, uint16_t *tile, int ts) {
    for (int y = -2; y < bh + 2; y++)
        for (int x = -2; x < bw + 2; x++) {
            const int yy = y0 + y, xx = x0 + x;
            tile[(y + 2) * ts + x + 2] = (yy >= 0 && yy < pl->h && xx >= 0 && xx < pl->w) ? (uint16_t)px(pl, yy, xx) : VERY_LARGE;
        }
}

/* dist_8x8_8bit_c / dist_8x8_16bit_c (Encoder/Codec/EbEncCdef.c:20-33,75-98): perceptual 8x8 distortion in
 * double arithmetic. Every operation is a correctly rounded IEEE double op, in this order. */
static uint64_t dist_8x8(const int *src /*filtered 64*/, const int *dst /*source picture 64*/, int coeff_shift) {
    uint64_t sum_s = 0, sum_d = 0, sum_s2 = 0, sum_d2 = 0, sum_sd = 0;
    for (int i = 0; i < 64; i++) {
        sum_s += src[i];
        sum_d += dst[i];
        sum_s2 += (uint64_t)(src[i] * src[i]);
        sum_d2 += (uint64_t)(dst[i] * dst[i]);
        sum_sd += (uint64_t)(src[i] * dst[i]);
    }
    const uint64_t svar = sum_s2 - ((sum_s * sum_s + 32) >> 6);
    const uint64_t dvar = sum_d2 - ((sum_d * sum_d + 32) >> 6);
    const double a = (double)(sum_d2 + sum_s2 - 2 * sum_sd) * .5;
    const double b = a * (double)(svar + dvar + (uint64_t)(400 << 2 * coeff_shift));
    const double c = sqrt((double)(20000 << 4 * coeff_shift) + (double)svar * (double)dvar);
    return (uint64_t)floor(.5 + b / c);
}

static int is_skip8(const uint8_t *skip8, int stride, int r8, int c8) { return skip8[r8 * stride + c8]; }

/* cdef_seg_search / cdef_seg_search16bit over the whole picture (EbCdefProcess.c:80-475) */
void orc_cdef_search(const SvtB200CdefSearchParams *p, const SvtB200Frame *recon, const SvtB200Frame *source,
                     const uint8_t *skip8, int32_t skip_stride, uint64_t *mse) {
    const int nvfb = (p->mi_rows + 15) / 16, nhfb = (p->mi_cols + 15) / 16;
    const int coeff_shift = recon->bit_depth > 8 ? recon->bit_depth - 8 : 0;
    memset(mse, 0, sizeof(uint64_t) * 2 * nvfb * nhfb * 64);
    uint16_t tile[68 * 72];
    for (int fbr = 0; fbr < nvfb; fbr++)
        for (int fbc = 0; fbc < nhfb; fbc++) {
            const int nvb = (p->mi_rows - 16 * fbr < 16 ? p->mi_rows - 16 * fbr : 16);
            const int nhb = (p->mi_cols - 16 * fbc < 16 ? p->mi_cols - 16 * fbc : 16);
            /* dlist: non-skip 8x8 blocks (svt_sb_compute_cdef_list); all-skip blocks are not searched */
            int by[64], bx[64], count = 0;
            for (int r = 0; r < nvb; r += 2)
                for (int c = 0; c < nhb; c += 2)
                    if (!is_skip8(skip8, skip_stride, (16 * fbr + r) >> 1, (16 * fbc + c) >> 1)) {
                        by[count] = r >> 1;
                        bx[count] = c >> 1;
                        count++;
                    }
            if (!count) continue;
            int dir[8][8], var[8][8];
            memset(dir, 0, sizeof(dir));
            memset(var, 0, sizeof(var));
            for (int pli = 0; pli < 3; pli++) {
                const PlaneRO rp = plane_of(recon, pli, p->mi_rows, p->mi_cols);
                const PlaneRO sp = plane_of(source, pli, p->mi_rows, p->mi_cols);
                const int sh = pli ? 1 : 0, bs = 8 >> sh; /* block size in this plane */
                const int bh = (nvb * 4) >> sh, bw = (nhb * 4) >> sh;
                load_tile(&rp, (fbr * 64) >> sh, (fbc * 64) >> sh, bh, bw, tile, 72);
                const uint16_t *in = tile + 2 * 72 + 2;
                if (pli == 0)
                    for (int b = 0; b < count; b++)
                        dir[by[b]][bx[b]] = orc_cdef_find_dir(in + 8 * by[b] * 72 + 8 * bx[b], 72, &var[by[b]][bx[b]], coeff_shift);
                const int damping = p->pri_damping + coeff_shift - (pli != 0);
                for (int gi = 0; gi < p->n_strengths; gi++) {
                    const int pri = p->pri_strength[gi] << coeff_shift, sec = p->sec_strength[gi] << coeff_shift;
                    uint64_t sum = 0;
                    for (int b = 0; b < count; b++) {
                        const int t = pli ? pri : adjust_strength(pri, var[by[b]][bx[b]]);
                        const int d = pri ? dir[by[b]][bx[b]] : 0;
                        int flt[64], org[64];
                        for (int i = 0; i < bs; i++)
                            for (int j = 0; j < bs; j++) {
                                const uint16_t *q = in + (by[b] * bs + i) * 72 + bx[b] * bs + j;
                                /* pri == sec == 0 takes the reference's copy path; the filter is the identity there */
                                flt[i * bs + j] = (pri == 0 && sec == 0) ? q[0] : cdef_sample(q, 72, t, sec, d, damping, damping, coeff_shift);
                                org[i * bs + j] = px(&sp, ((fbr * 64) >> sh) + by[b] * bs + i, ((fbc * 64) >> sh) + bx[b] * bs + j);
                            }
                        if (pli == 0)
                            sum += dist_8x8(flt, org, coeff_shift);
                        else
                            for (int i = 0; i < bs * bs; i++) sum += (uint64_t)((org[i] - flt[i]) * (org[i] - flt[i]));
                    }
                    sum >>= 2 * coeff_shift;
                    uint64_t *m = mse + ((size_t)(pli ? 1 : 0) * nvfb * nhfb + fbr * nhfb + fbc) * 64 + gi;
                    if (pli < 2)
                        *m = sum;
                    else
                        *m += sum;
                }
            }
        }
}

static void put_px(const SvtB200Frame *f, int pli, int y, int x, int v) {
    void *p = pli == 0 ? f->y : pli == 1 ? f->cb : f->cr;
    const int stride = pli == 0 ? f->stride_y : f->stride_c;
    if (f->bit_depth > 8)
        ((uint16_t *)p)[(size_t)y * stride + x] = (uint16_t)v;
    else
        ((uint8_t *)p)[(size_t)y * stride + x] = (uint8_t)v;
}

/* svt_av1_cdef_frame / av1_cdef_frame16bit (Encoder/Codec/EbEncCdef.c:292-1030), out of place: the reference
 * filters in place but feeds every block from saved PRE-filter rows/columns (linebuf/colbuf), which is the
 * same function of the pre-CDEF picture. */
void orc_cdef_apply(const SvtB200CdefApplyParams *p, const SvtB200Frame *recon, const SvtB200Frame *out,
                    const uint8_t *skip8, int32_t skip_stride, const int8_t *fb_strength_idx) {
    const int nvfb = (p->mi_rows + 15) / 16, nhfb = (p->mi_cols + 15) / 16;
    const int coeff_shift = recon->bit_depth > 8 ? recon->bit_depth - 8 : 0;
    uint16_t tile[68 * 72];
    for (int pli = 0; pli < 3; pli++) { /* start from a copy */
        const PlaneRO rp = plane_of(recon, pli, p->mi_rows, p->mi_cols);
        for (int y = 0; y < rp.h; y++)
            for (int x = 0; x < rp.w; x++) put_px(out, pli, y, x, px(&rp, y, x));
    }
    for (int fbr = 0; fbr < nvfb; fbr++)
        for (int fbc = 0; fbc < nhfb; fbc++) {
            const int idx = fb_strength_idx[fbr * nhfb + fbc];
            if (idx < 0) continue;
            const int nvb = (p->mi_rows - 16 * fbr < 16 ? p->mi_rows - 16 * fbr : 16);
            const int nhb = (p->mi_cols - 16 * fbc < 16 ? p->mi_cols - 16 * fbc : 16);
            int level = p->y_strength[idx] / 4, sec = p->y_strength[idx] % 4;
            sec += sec == 3;
            int uv_level = p->uv_strength[idx] / 4, uv_sec = p->uv_strength[idx] % 4;
            uv_sec += uv_sec == 3;
            if (level == 0 && sec == 0 && uv_level == 0 && uv_sec == 0) continue;
            int by[64], bx[64], count = 0;
            for (int r = 0; r < nvb; r += 2)
                for (int c = 0; c < nhb; c += 2)
                    if (!is_skip8(skip8, skip_stride, (16 * fbr + r) >> 1, (16 * fbc + c) >> 1)) {
                        by[count] = r >> 1;
                        bx[count] = c >> 1;
                        count++;
                    }
            if (!count) continue;
            int dir[8][8], var[8][8];
            memset(dir, 0, sizeof(dir));
            memset(var, 0, sizeof(var));
            for (int pli = 0; pli < 3; pli++) {
                const PlaneRO rp = plane_of(recon, pli, p->mi_rows, p->mi_cols);
                const int sh = pli ? 1 : 0, bs = 8 >> sh;
                const int bh = (nvb * 4) >> sh, bw = (nhb * 4) >> sh;
                load_tile(&rp, (fbr * 64) >> sh, (fbc * 64) >> sh, bh, bw, tile, 72);
                const uint16_t *in = tile + 2 * 72 + 2;
                if (pli == 0) /* svt_cdef_filter_fb: directions are found on luma whatever the strength */
                    for (int b = 0; b < count; b++)
                        dir[by[b]][bx[b]] = orc_cdef_find_dir(in + 8 * by[b] * 72 + 8 * bx[b], 72, &var[by[b]][bx[b]], coeff_shift);
                const int pri = (pli ? uv_level : level) << coeff_shift, s2 = (pli ? uv_sec : sec) << coeff_shift;
                const int damping = p->damping + coeff_shift - (pli != 0);
                for (int b = 0; b < count; b++) {
                    const int t = pli ? pri : adjust_strength(pri, var[by[b]][bx[b]]);
                    const int d = pri ? dir[by[b]][bx[b]] : 0;
                    for (int i = 0; i < bs; i++)
                        for (int j = 0; j < bs; j++) {
                            const uint16_t *q = in + (by[b] * bs + i) * 72 + bx[b] * bs + j;
                            put_px(out, pli, ((fbr * 64) >> sh) + by[b] * bs + i, ((fbc * 64) >> sh) + bx[b] * bs + j,
                                   cdef_sample(q, 72, t, s2, d, damping, damping, coeff_shift));
                        }
                }
            }
        }
}

/* get_cdef_filter_strengths + the `sec_strength + (sec_strength == 3)` of the callers */
int orc_cdef_strength_table(int pick_method, SvtB200CdefSearchParams *p) {
    static const int n[4] = {64, 32, 20, 10};
    static const int pri1[8] = {0, 1, 2, 3, 5, 7, 10, 13}, pri2[5] = {0, 2, 4, 8, 14}, sec3[2] = {0, 2};
    const int tot_sec = pick_method == 3 ? 2 : 4;
    p->n_strengths = n[pick_method];
    for (int gi = 0; gi < p->n_strengths; gi++) {
        int pi = gi / tot_sec, si = gi % tot_sec;
        int pri = pick_method == 0 ? pi : pick_method == 1 ? pri1[pi] : pri2[pi];
        int sec = pick_method == 3 ? sec3[si] : si;
        p->pri_strength[gi] = pri;
        p->sec_strength[gi] = sec + (sec == 3);
    }
    return p->n_strengths;
}

/* ---- the strength decision: finish_cdef_search (EbEncCdef.c:1167-1340) ------------------------------------------------
 * sb list = the filter blocks that are not all-skip (svt_sb_all_skip, :1243), in raster order; mse[0] luma, mse[1] chroma
 * sums per strength index.  svt_search_one_dual_c (:1070-1114): given nb already chosen (luma, chroma) pairs, add the pair
 * (j, k) that minimises sum over blocks of min(best so far, mse0[j] + mse1[k]); first minimum in (j outer, k inner) order.
 * joint_strength_search_dual (:1136-1160): greedy for nb = 1..n, then 4 n refinement rounds that drop the oldest pair and
 * re-add.  The number of signalled sets (1, 2, 4, 8) minimises RDCOST(lambda, cost_literal(sb_count * bits + n * 12), 16 * mse)
 * with strict improvement (:1263-1283); every block then takes its best set (:1287-1298). */
static uint64_t one_dual(int *lev0, int *lev1, int nb, const uint64_t *mse0, const uint64_t *mse1, const int *sb, int sb_count, int n) {
    static uint64_t tot[64][64];
    uint64_t best_tot = (uint64_t)1 << 63;
    int b0 = 0, b1 = 0;
    memset(tot, 0, sizeof(tot));
    for (int i = 0; i < sb_count; i++) {
        const uint64_t *m0 = mse0 + (size_t)sb[i] * 64, *m1 = mse1 + (size_t)sb[i] * 64;
        uint64_t best = (uint64_t)1 << 63;
        for (int g = 0; g < nb; g++) {
            const uint64_t c = m0[lev0[g]] + m1[lev1[g]];
            if (c < best) best = c;
        }
        for (int j = 0; j < n; j++)
            for (int k = 0; k < n; k++) {
                const uint64_t c = m0[j] + m1[k];
                tot[j][k] += c < best ? c : best;
            }
    }
    for (int j = 0; j < n; j++)
        for (int k = 0; k < n; k++)
            if (tot[j][k] < best_tot) best_tot = tot[j][k], b0 = j, b1 = k;
    lev0[nb] = b0, lev1[nb] = b1;
    return best_tot;
}

void orc_cdef_decide(const SvtB200CdefDecideParams *p, const uint64_t *mse, const uint8_t *skip8, int skip_stride,
                     SvtB200CdefDecision *out, int8_t *fb_strength_idx) {
    const int nvfb = (p->mi_rows + 15) / 16, nhfb = (p->mi_cols + 15) / 16, nfb = nvfb * nhfb;
    const int rows8 = (p->mi_rows + 1) / 2, cols8 = (p->mi_cols + 1) / 2;
    const uint64_t *mse0 = mse, *mse1 = mse + (size_t)nfb * 64;
    int *sb = malloc(sizeof(int) * (size_t)nfb), sb_count = 0;
    for (int fb = 0; fb < nfb; fb++) {
        const int fbr = fb / nhfb, fbc = fb % nhfb;
        int all_skip = 1;
        for (int r = 8 * fbr; r < 8 * fbr + 8 && r < rows8; r++)
            for (int c = 8 * fbc; c < 8 * fbc + 8 && c < cols8; c++) all_skip &= skip8[(size_t)r * skip_stride + c] != 0;
        fb_strength_idx[fb] = -1;
        if (!all_skip) sb[sb_count++] = fb;
    }
    memset(out, 0, sizeof(*out));
    uint64_t best_cost = (uint64_t)1 << 63;
    for (int bits = 0; bits <= 3; bits++) {
        int lev0[8], lev1[8] = {0};
        const int n = 1 << bits;
        uint64_t tot = (uint64_t)1 << 63;
        for (int i = 0; i < n; i++) tot = one_dual(lev0, lev1, i, mse0, mse1, sb, sb_count, p->n_strengths);
        for (int i = 0; i < 4 * n; i++) {
            for (int j = 0; j < n - 1; j++) lev0[j] = lev0[j + 1], lev1[j] = lev1[j + 1];
            tot = one_dual(lev0, lev1, n - 1, mse0, mse1, sb, sb_count, p->n_strengths);
        }
        const int total_bits = sb_count * bits + n * 6 * 2;
        const uint64_t rate = (uint64_t)((int64_t)total_bits * 512), dist = tot * 16;
        const uint64_t cost = ((rate * p->lambda + 256) >> 9) + dist * 128;
        if (cost < best_cost) {
            best_cost = cost;
            out->cdef_bits = bits;
            for (int j = 0; j < n; j++) out->y_index[j] = lev0[j], out->uv_index[j] = lev1[j];
        }
    }
    out->nb_cdef_strengths = 1 << out->cdef_bits;
    out->sb_count = sb_count;
    for (int i = 0; i < sb_count; i++) {
        const uint64_t *m0 = mse0 + (size_t)sb[i] * 64, *m1 = mse1 + (size_t)sb[i] * 64;
        uint64_t best = (uint64_t)1 << 63;
        int bg = 0;
        for (int g = 0; g < out->nb_cdef_strengths; g++) {
            const uint64_t c = m0[out->y_index[g]] + m1[out->uv_index[g]];
            if (c < best) best = c, bg = g;
        }
        fb_strength_idx[sb[i]] = (int8_t)bg;
    }
    for (int j = 0; j < out->nb_cdef_strengths; j++) {
        out->y_strength[j] = p->filter_strength[out->y_index[j]];
        out->uv_strength[j] = p->filter_strength[out->uv_index[j]];
    }
    free(sb);
}

/* nb_cdef_strengths[pick_method] and STORE_CDEF_FILTER_STRENGTH per index (get_cdef_filter_strengths, EbDefinitions.h:1696) */
int orc_cdef_decide_table(int pick_method, SvtB200CdefDecideParams *p) {
    SvtB200CdefSearchParams sp;
    const int n = orc_cdef_strength_table(pick_method, &sp);
    if (n <= 0) return n;
    p->n_strengths = n;
    for (int g = 0; g < n; g++) {
        const int sec = sp.sec_strength[g] == 4 ? 3 : sp.sec_strength[g]; /* the table stores sec + (sec == 3) */
        p->filter_strength[g] = pick_method == 0 ? g : sp.pri_strength[g] * 4 + sec;
    }
    return n;
}
