/*
 * interp_oracle.c — CPU restatement of the translational inter-prediction path.
 *
 * TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this file; the
 * product (libsvtav1_b200.so) never links or calls it.
 *
 * Follows, in svt-av1 v0.8.6:
 *   Source/Lib/Common/Codec/EbInterPrediction.c
 *     :349-393  svt_av1_convolve_2d_sr_c        :395-423  svt_av1_convolve_y_sr_c
 *     :425-453  svt_av1_convolve_x_sr_c         :455-470  svt_av1_convolve_2d_copy_sr_c
 *     :552-610  svt_av1_jnt_convolve_2d_c       :612-656  svt_av1_jnt_convolve_y_c
 *     :658-702  svt_av1_jnt_convolve_x_c        :704-745  svt_av1_jnt_convolve_2d_copy_c
 *     :747-1145 the highbd forms (same arithmetic with bd in the offsets and the final clip)
 *     :1251-1270 av1_get_interp_filter_params_with_block_size / av1_get_convolve_filter_params
 *     :1368-1431 svt_inter_predictor (dispatch on subpel_x != 0, subpel_y != 0, is_compound)
 *   Source/Lib/Common/Codec/convolve.h:44-71   get_conv_params_no_round
 *   Source/Lib/Common/Codec/convolve.c:249-308 svt_aom_convolve8_horiz_c / _vert_c
 *   Source/Lib/Encoder/Codec/EbEncInterPrediction.c
 *     :24-45     clamp_mv_to_umv_border_sb      :3591-3661 compute_subpel_params (unscaled branch)
 *     :3663-3762 enc_make_inter_predictor
 *
 * The sixteen reference functions differ only in which of the two 8-tap passes runs and in how the result is rounded,
 * so they are restated as ONE routine parameterised by (filter x?, filter y?, compound?); tests/test_oracle_interp.py
 * pins it against each of the sixteen reference functions and against enc_make_inter_predictor (oracle/_ref).
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

/* AV1 interpolation kernels (spec section 7.11.3.4 "Block inter prediction process", Subpel_Filters); rows = 1/16 positions.
 * Pinned against the reference's sub_pel_filters_* / bilinear_filters symbols by the tests. */
static const int8_t k_regular[16][8] = {{0, 0, 0, 64, 0, 0, 0, 0},     {0, 1, -3, 63, 4, -1, 0, 0},   {0, 1, -5, 61, 9, -2, 0, 0},
                                        {0, 1, -6, 58, 14, -4, 1, 0},  {0, 1, -7, 55, 19, -5, 1, 0},  {0, 1, -7, 51, 24, -6, 1, 0},
                                        {0, 1, -8, 47, 29, -6, 1, 0},  {0, 1, -7, 42, 33, -6, 1, 0},  {0, 1, -7, 38, 38, -7, 1, 0},
                                        {0, 1, -6, 33, 42, -7, 1, 0},  {0, 1, -6, 29, 47, -8, 1, 0},  {0, 1, -6, 24, 51, -7, 1, 0},
                                        {0, 1, -5, 19, 55, -7, 1, 0},  {0, 1, -4, 14, 58, -6, 1, 0},  {0, 0, -2, 9, 61, -5, 1, 0},
                                        {0, 0, -1, 4, 63, -3, 1, 0}};
static const int8_t k_smooth[16][8] = {{0, 0, 0, 64, 0, 0, 0, 0},    {0, 1, 14, 31, 17, 1, 0, 0},  {0, 0, 13, 31, 18, 2, 0, 0},
                                       {0, 0, 11, 31, 20, 2, 0, 0},  {0, 0, 10, 30, 21, 3, 0, 0},  {0, 0, 9, 29, 22, 4, 0, 0},
                                       {0, 0, 8, 28, 23, 5, 0, 0},   {0, -1, 8, 27, 24, 6, 0, 0},  {0, -1, 7, 26, 26, 7, -1, 0},
                                       {0, 0, 6, 24, 27, 8, -1, 0},  {0, 0, 5, 23, 28, 8, 0, 0},   {0, 0, 4, 22, 29, 9, 0, 0},
                                       {0, 0, 3, 21, 30, 10, 0, 0},  {0, 0, 2, 20, 31, 11, 0, 0},  {0, 0, 2, 18, 31, 13, 0, 0},
                                       {0, 0, 1, 17, 31, 14, 1, 0}};
static const int8_t k_sharp[16][8] = {{0, 0, 0, 64, 0, 0, 0, 0},       {-1, 1, -3, 63, 4, -1, 1, 0},    {-1, 3, -6, 62, 8, -3, 2, -1},
                                      {-1, 4, -9, 60, 13, -5, 3, -1},  {-2, 5, -11, 58, 19, -7, 3, -1}, {-2, 5, -11, 54, 24, -9, 4, -1},
                                      {-2, 5, -12, 50, 30, -10, 4, -1}, {-2, 5, -12, 45, 35, -11, 5, -1}, {-2, 6, -12, 40, 40, -12, 6, -2},
                                      {-1, 5, -11, 35, 45, -12, 5, -2}, {-1, 4, -10, 30, 50, -12, 5, -2}, {-1, 4, -9, 24, 54, -11, 5, -2},
                                      {-1, 3, -7, 19, 58, -11, 5, -2}, {-1, 3, -5, 13, 60, -9, 4, -1},  {-1, 2, -3, 8, 62, -6, 3, -1},
                                      {0, 1, -1, 4, 63, -3, 1, -1}};
static const int8_t k_regular4[16][8] = {{0, 0, 0, 64, 0, 0, 0, 0},   {0, 0, -2, 63, 4, -1, 0, 0},  {0, 0, -4, 61, 9, -2, 0, 0},
                                         {0, 0, -5, 58, 14, -3, 0, 0}, {0, 0, -6, 55, 19, -4, 0, 0}, {0, 0, -6, 51, 24, -5, 0, 0},
                                         {0, 0, -7, 47, 29, -5, 0, 0}, {0, 0, -6, 42, 33, -5, 0, 0}, {0, 0, -6, 38, 38, -6, 0, 0},
                                         {0, 0, -5, 33, 42, -6, 0, 0}, {0, 0, -5, 29, 47, -7, 0, 0}, {0, 0, -5, 24, 51, -6, 0, 0},
                                         {0, 0, -4, 19, 55, -6, 0, 0}, {0, 0, -3, 14, 58, -5, 0, 0}, {0, 0, -2, 9, 61, -4, 0, 0},
                                         {0, 0, -1, 4, 63, -2, 0, 0}};
static const int8_t k_smooth4[16][8] = {{0, 0, 0, 64, 0, 0, 0, 0},   {0, 0, 15, 31, 17, 1, 0, 0}, {0, 0, 13, 31, 18, 2, 0, 0},
                                        {0, 0, 11, 31, 20, 2, 0, 0}, {0, 0, 10, 30, 21, 3, 0, 0}, {0, 0, 9, 29, 22, 4, 0, 0},
                                        {0, 0, 8, 28, 23, 5, 0, 0},  {0, 0, 7, 27, 24, 6, 0, 0},  {0, 0, 6, 26, 26, 6, 0, 0},
                                        {0, 0, 6, 24, 27, 7, 0, 0},  {0, 0, 5, 23, 28, 8, 0, 0},  {0, 0, 4, 22, 29, 9, 0, 0},
                                        {0, 0, 3, 21, 30, 10, 0, 0}, {0, 0, 2, 20, 31, 11, 0, 0}, {0, 0, 2, 18, 31, 13, 0, 0},
                                        {0, 0, 1, 17, 31, 15, 0, 0}};

/* av1_get_interp_filter_params_with_block_size (:1251-1262) + av1_get_interp_filter_subpel_kernel; the tables above hold
 * half the coefficient (every AV1 tap is even), bilinear is computed. */
ORC_API void orc_interp_kernel(int filter, int w, int subpel, int16_t out[8]) {
    const int8_t(*t)[8] = NULL;
    subpel &= 15;
    if (filter == 3) { /* BILINEAR: {128 - 8 s, 8 s} on taps 3, 4 */
        memset(out, 0, 8 * sizeof(int16_t));
        out[3] = (int16_t)(128 - 8 * subpel);
        out[4] = (int16_t)(8 * subpel);
        return;
    }
    if (w <= 4)
        t = filter == 1 ? k_smooth4 : k_regular4; /* sharp falls back to the regular 4-tap kernel */
    else
        t = filter == 0 ? k_regular : filter == 1 ? k_smooth : k_sharp;
    for (int k = 0; k < 8; k++) out[k] = (int16_t)(2 * t[subpel][k]);
}

static inline int rshift_round(int v, int n) { return (v + ((1 << n) >> 1)) >> n; } /* ROUND_POWER_OF_TWO */
static inline int clip_bd(int v, int bd) {
    const int mx = (1 << bd) - 1;
    return v < 0 ? 0 : v > mx ? mx : v;
}
static inline int sample(const void *p, int hbd, ptrdiff_t i) { return hbd ? ((const uint16_t *)p)[i] : ((const uint8_t *)p)[i]; }
static inline void put(void *p, int hbd, ptrdiff_t i, int v) {
    if (hbd)
        ((uint16_t *)p)[i] = (uint16_t)v;
    else
        ((uint8_t *)p)[i] = (uint8_t)v;
}

/* One call of convolve[sx][sy][is_compound] / convolveHbd[..] (:1147-1175). src/dst address the block's sample (0,0).
 * fx / fy: the 8 taps (NULL = that pass is not run, i.e. the _copy / _y / _x variants).
 * compound == 0: dst <- prediction.  compound, do_average == 0: conv_dst <- intermediate.  compound, do_average == 1:
 * dst <- average of conv_dst and this reference's intermediate (plain or distance weighted). */
ORC_API void orc_convolve(const void *src, int hbd, int src_stride, void *dst, int dst_stride, int w, int h, const int16_t *fx,
                          const int16_t *fy, int round_0, int round_1, int bd, int compound, int do_average, int use_jnt,
                          int fwd_offset, int bck_offset, uint16_t *conv_dst, int conv_stride) {
    const int offset_bits = bd + 14 - round_0;
    const int round_offset = (1 << (offset_bits - round_1)) + (1 << (offset_bits - round_1 - 1));
    const int bits2 = 14 - round_0 - round_1; /* "bits" of the 2d / copy forms, "round_bits" of the jnt forms */
    int16_t *im = NULL;
    if (fx && fy) { /* first pass of the 2-D forms: h + 7 rows, offset 1 << (bd + 6), rounded by round_0, kept as int16 */
        im = (int16_t *)malloc(sizeof(int16_t) * (size_t)w * (h + 7));
        for (int y = 0; y < h + 7; y++)
            for (int x = 0; x < w; x++) {
                int sum = 1 << (bd + 6);
                for (int k = 0; k < 8; k++) sum += fx[k] * sample(src, hbd, (ptrdiff_t)(y - 3) * src_stride + x - 3 + k);
                im[y * w + x] = (int16_t)rshift_round(sum, round_0);
            }
    }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int res; /* non-compound: the prediction sample; compound: this reference's intermediate */
            if (fx && fy) {
                int sum = 1 << offset_bits;
                for (int k = 0; k < 8; k++) sum += fy[k] * im[(y + k) * w + x];
                res = rshift_round(sum, round_1);
                if (!compound) {
                    const int16_t r16 = (int16_t)(uint16_t)(res - round_offset);
                    res = clip_bd(rshift_round(r16, bits2), bd);
                } else
                    res = (uint16_t)res;
            } else if (fx) {
                int sum = 0;
                for (int k = 0; k < 8; k++) sum += fx[k] * sample(src, hbd, (ptrdiff_t)y * src_stride + x - 3 + k);
                sum = rshift_round(sum, round_0);
                res = compound ? (1 << (7 - round_1)) * sum + round_offset : clip_bd(rshift_round(sum, 7 - round_0), bd);
            } else if (fy) {
                int sum = 0;
                for (int k = 0; k < 8; k++) sum += fy[k] * sample(src, hbd, (ptrdiff_t)(y - 3 + k) * src_stride + x);
                res = compound ? rshift_round(sum * (1 << (7 - round_0)), round_1) + round_offset : clip_bd(rshift_round(sum, 7), bd);
            } else {
                const int p = sample(src, hbd, (ptrdiff_t)y * src_stride + x);
                res = compound ? (uint16_t)((uint16_t)(p << bits2) + (uint16_t)round_offset) : p;
            }
            if (!compound)
                put(dst, hbd, (ptrdiff_t)y * dst_stride + x, res);
            else if (!do_average)
                conv_dst[y * conv_stride + x] = (uint16_t)res;
            else {
                int tmp = conv_dst[y * conv_stride + x];
                tmp = use_jnt ? (tmp * fwd_offset + res * bck_offset) >> 4 : (tmp + res) >> 1;
                tmp -= round_offset;
                put(dst, hbd, (ptrdiff_t)y * dst_stride + x, clip_bd(rshift_round(tmp, bits2), bd));
            }
        }
    free(im);
}

/* svt_aom_convolve8_horiz_c / _vert_c (convolve.c:249-308): `table` = the [16][8] kernel table (get_filter_base),
 * q0 = the start phase (get_filter_offset), stepping `step` sixteenths per output sample. */
ORC_API void orc_convolve8(const uint8_t *src, ptrdiff_t src_stride, uint8_t *dst, ptrdiff_t dst_stride, const int16_t *table,
                           int q0, int step, int w, int h, int vert) {
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const int q = q0 + (vert ? y : x) * step;
            const int16_t *f = table + 8 * (q & 15);
            int sum = 0;
            for (int k = 0; k < 8; k++)
                sum += f[k] * (vert ? src[((q >> 4) - 3 + k) * src_stride + x] : src[y * src_stride + (q >> 4) - 3 + k]);
            dst[y * dst_stride + x] = (uint8_t)clip_bd(rshift_round(sum, 7), 8);
        }
}

/* get_conv_params_no_round (convolve.h:44-71) */
static void conv_rounds(int bd, int compound, int *r0, int *r1) {
    *r0 = 3;
    *r1 = compound ? 7 : 14 - *r0;
    const int intbufrange = bd + 7 - *r0 + 2;
    if (intbufrange > 16) {
        *r0 += intbufrange - 16;
        if (!compound) *r1 -= intbufrange - 16;
    }
}
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }

/* The jobs of svt_b200_inter_predict, one after the other: compute_subpel_params' unscaled branch (clamp the MV, split
 * it into whole-sample position and 1/16 phase), svt_inter_predictor's dispatch, and for n_refs == 2 the do_average
 * 0 / 1 pair through a CONV_BUF. */
ORC_API void orc_inter_predict(const SvtB200Frame *refs, int n_ref_frames, const SvtB200Frame *pred, const SvtB200InterJob *jobs,
                               int n_jobs) {
    const int bd = pred->bit_depth, hbd = bd > 8;
    uint16_t *conv = (uint16_t *)malloc(sizeof(uint16_t) * 128 * 128);
    (void)n_ref_frames;
    for (int j = 0; j < n_jobs; j++) {
        const SvtB200InterJob *b = &jobs[j];
        const int ss = b->plane != 0, compound = b->n_refs == 2;
        const int dstride = b->plane ? pred->stride_c : pred->stride_y;
        uint8_t *dplane = (uint8_t *)(b->plane == 0 ? pred->y : b->plane == 1 ? pred->cb : pred->cr);
        void *dst = dplane + (((ptrdiff_t)b->dst_y * dstride + b->dst_x) << hbd);
        int r0, r1;
        conv_rounds(bd, compound, &r0, &r1);
        for (int r = 0; r < b->n_refs; r++) {
            const SvtB200Frame *rf = &refs[b->ref[r]];
            const int sstride = b->plane ? rf->stride_c : rf->stride_y;
            const uint8_t *splane = (const uint8_t *)(b->plane == 0 ? rf->y : b->plane == 1 ? rf->cb : rf->cr);
            /* clamp_mv_to_umv_border_sb: MV in 1/16 sample of this plane */
            const int spel_left = (4 + b->bw) << 4, spel_right = spel_left - 16;
            const int spel_top = (4 + b->bh) << 4, spel_bottom = spel_top - 16;
            const int sc = 1 << (1 - ss);
            int row = (int16_t)(b->mv_row[r] * sc), col = (int16_t)(b->mv_col[r] * sc);
            col = (int16_t)clampi(col, b->mb_to_left_edge * sc - spel_left, b->mb_to_right_edge * sc + spel_right);
            row = (int16_t)clampi(row, b->mb_to_top_edge * sc - spel_top, b->mb_to_bottom_edge * sc + spel_bottom);
            const int sx = col & 15, sy = row & 15;
            const int px = ((b->pre_x << 4) + col) >> 4, py = ((b->pre_y << 4) + row) >> 4;
            int16_t fx[8], fy[8];
            orc_interp_kernel(b->filter_x, b->bw, sx, fx);
            orc_interp_kernel(b->filter_y, b->bh, sy, fy);
            orc_convolve(splane + (((ptrdiff_t)py * sstride + px) << hbd), hbd, sstride, dst, dstride, b->bw, b->bh, sx ? fx : NULL,
                         sy ? fy : NULL, r0, r1, bd, compound, r, b->use_jnt_comp_avg, b->fwd_offset, b->bck_offset, conv, 128);
        }
    }
    free(conv);
}
