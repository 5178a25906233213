/* pa_oracle.c - CPU restatement of the picture-analysis block statistics (TEST INFRASTRUCTURE).
 *
 * Follows Source/Lib/Encoder/Codec/EbPictureAnalysisProcess.c:
 *   compute_block_mean_compute_variance   :1005-2575   (per 64x64 SB: 85 luma means and 85 variances, EbMeTierZeroPu order:
 *                                                       64x64, 4 x 32x32, 16 x 16x16, 64 x 8x8 in raster order per level)
 *   svt_compute_sub_mean_8x8_c / compute_sub_mean_squared_values_c / svt_compute_interm_var_four8x8_c   :310-378
 *   svt_compute_mean_squared_values_c :287-308 and compute_mean_8x8 (the BLOCK_MEAN_PREC_FULL flavour)
 *   compute_chroma_block_mean :493-1003, zero_out_chroma_block_mean :432-487
 *   compute_picture_spatial_statistics :2929-2974 (pic_avg_variance)
 * 8x8 level: mean = (sum over the rows 0,2,4,6) << 3 and mean of squares = (sum of squares over the same rows) << 11
 * (BLOCK_MEAN_PREC_SUB, the sequence default, EbSequenceControlSet.c:192), or (sum << 8) / 64 and (sum of squares << 16) / 64
 * over all rows (FULL); each larger level is (a + b + c + d) >> 2 of both; mean out = level mean >> 8; variance out =
 * (uint16)((mean of squares - mean * mean) >> 16) in uint64 arithmetic. */
#include <stdint.h>
#include <string.h>

#include "oracle.h"

static void level_up(const uint64_t *in, int n_in_side, uint64_t *out) {
    const int n = n_in_side / 2;
    for (int r = 0; r < n; r++)
        for (int c = 0; c < n; c++)
            out[r * n + c] = (in[(2 * r) * n_in_side + 2 * c] + in[(2 * r) * n_in_side + 2 * c + 1] + in[(2 * r + 1) * n_in_side + 2 * c] +
                              in[(2 * r + 1) * n_in_side + 2 * c + 1]) >> 2;
}

/* luma: y points at the SB's top-left sample inside a padded plane (64x64 samples are read) */
void orc_sb_mean_variance(const uint8_t *y, int stride, int full_precision, uint8_t mean_out[85], uint16_t var_out[85]) {
    uint64_t m8[64], s8[64], m16[16], s16[16], m32[4], s32[4], m64, s64;
    for (int b = 0; b < 64; b++) {
        const uint8_t *p = y + (b / 8) * 8 * stride + (b % 8) * 8;
        uint64_t sum = 0, sq = 0;
        for (int r = 0; r < 8; r += full_precision ? 1 : 2)
            for (int c = 0; c < 8; c++) {
                sum += p[r * stride + c];
                sq += (uint64_t)p[r * stride + c] * p[r * stride + c];
            }
        m8[b] = full_precision ? (sum << 8) / 64 : sum << 3;
        s8[b] = full_precision ? (sq << 16) / 64 : sq << 11;
    }
    level_up(m8, 8, m16), level_up(s8, 8, s16);
    level_up(m16, 4, m32), level_up(s16, 4, s32);
    level_up(m32, 2, &m64), level_up(s32, 2, &s64);
    mean_out[0] = (uint8_t)(m64 >> 8), var_out[0] = (uint16_t)((s64 - m64 * m64) >> 16);
    for (int i = 0; i < 4; i++) mean_out[1 + i] = (uint8_t)(m32[i] >> 8), var_out[1 + i] = (uint16_t)((s32[i] - m32[i] * m32[i]) >> 16);
    for (int i = 0; i < 16; i++) mean_out[5 + i] = (uint8_t)(m16[i] >> 8), var_out[5 + i] = (uint16_t)((s16[i] - m16[i] * m16[i]) >> 16);
    for (int i = 0; i < 64; i++) mean_out[21 + i] = (uint8_t)(m8[i] >> 8), var_out[21 + i] = (uint16_t)((s8[i] - m8[i] * m8[i]) >> 16);
}

/* chroma of a COMPLETE SB: c points at the SB's 32x32 chroma samples; writes entries 0..20 of mean_out (64x64, 32x32,
 * 16x16 levels); the 64x64 value uses the reference's expression (b[0] + b[1] + b[3] + b[3]) >> 2 (:900-905). */
void orc_sb_chroma_mean(const uint8_t *c, int stride, int full_precision, uint8_t mean_out[21]) {
    uint64_t m16[16], m32[4];
    for (int b = 0; b < 16; b++) {
        const uint8_t *p = c + (b / 4) * 8 * stride + (b % 4) * 8;
        uint64_t sum = 0;
        for (int r = 0; r < 8; r += full_precision ? 1 : 2)
            for (int x = 0; x < 8; x++) sum += p[r * stride + x];
        m16[b] = full_precision ? (sum << 8) / 64 : sum << 3;
    }
    level_up(m16, 4, m32);
    const uint64_t m64 = (m32[0] + m32[1] + m32[3] + m32[3]) >> 2;
    mean_out[0] = (uint8_t)(m64 >> 8);
    for (int i = 0; i < 4; i++) mean_out[1 + i] = (uint8_t)(m32[i] >> 8);
    for (int i = 0; i < 16; i++) mean_out[5 + i] = (uint8_t)(m16[i] >> 8);
}

/* whole picture (compute_picture_spatial_statistics): SBs in raster order; planes point at sample (0,0) and must be
 * readable up to the next multiple of 64 (the reference's padded input picture).  Outputs [n_sb][85] / [n_sb][21]
 * (chroma zero for incomplete SBs).  Returns pic_avg_variance. */
uint16_t orc_picture_mean_variance(const uint8_t *y, int stride_y, const uint8_t *cb, const uint8_t *cr, int stride_c, int width, int height,
                                   int full_precision, uint8_t *y_mean, uint16_t *variance, uint8_t *cb_mean, uint8_t *cr_mean) {
    const int sbw = (width + 63) / 64, sbh = (height + 63) / 64;
    uint64_t tot = 0;
    for (int sy = 0; sy < sbh; sy++)
        for (int sx = 0; sx < sbw; sx++) {
            const int sb = sy * sbw + sx;
            orc_sb_mean_variance(y + (size_t)sy * 64 * stride_y + sx * 64, stride_y, full_precision, y_mean + sb * 85, variance + sb * 85);
            const int complete = sx * 64 + 64 <= width && sy * 64 + 64 <= height;
            if (cb_mean && cr_mean) {
                if (complete) {
                    orc_sb_chroma_mean(cb + (size_t)sy * 32 * stride_c + sx * 32, stride_c, full_precision, cb_mean + sb * 21);
                    orc_sb_chroma_mean(cr + (size_t)sy * 32 * stride_c + sx * 32, stride_c, full_precision, cr_mean + sb * 21);
                } else {
                    memset(cb_mean + sb * 21, 0, 21);
                    memset(cr_mean + sb * 21, 0, 21);
                }
            }
            tot += variance[sb * 85];
        }
    return (uint16_t)(tot / (uint64_t)(sbw * sbh));
}

/* ---- open-loop intra search, the flavour of presets >= 5 (TPL "opt" controls) -------------------------------------
 * open_loop_intra_search_mb (EbMotionEstimation.c:3043-3155) with tpl_ctrls.tpl_opt_flag = 1 (set_tpl_controls,
 * EbPictureDecisionProcess.c:3899-3935: levels 5 and up): the mode loop runs DC_PRED only.  Per 16x16 macroblock whose
 * origin is inside the picture: neighbours from the SOURCE (update_neighbor_samples_array_open_loop_mb,
 * EbEncIntraPrediction.c:1201-1280: above = the row over the block, at most min(32, width - x) samples, the rest 127;
 * left = the column left of it, at most min(32, height - y) samples, the rest 129), dc_pred[x > 0][y > 0][TX_16X16]
 * (EbIntraPrediction.c:2606-2631), residual, svt_av1_wht_fwd_txfm = the forward 16x16 DCT_DCT (EbTransforms.c:3827),
 * intra_cost = svt_aom_satd = sum of |coefficient|.  y points at sample (0,0) of a plane that is readable (padded, as the
 * reference's input picture is) up to the next multiple of 16 in both directions.  cost: [mb rows][mb cols]. */
void orc_ois_dc_picture(const uint8_t *y, int stride, int width, int height, int64_t *cost) {
    const int mbw = (width + 15) / 16, mbh = (height + 15) / 16;
    for (int my = 0; my < mbh; my++)
        for (int mx = 0; mx < mbw; mx++) {
            const int x = mx * 16, yy = my * 16;
            int above[16], left[16];
            const int na = width - x < 32 ? width - x : 32, nl = height - yy < 32 ? height - yy : 32;
            for (int i = 0; i < 16; i++) {
                above[i] = (yy > 0 && i < na) ? y[(size_t)(yy - 1) * stride + x + i] : 127;
                left[i] = (x > 0 && i < nl) ? y[(size_t)(yy + i) * stride + x - 1] : 129;
            }
            int sa = 0, sl = 0, dc;
            for (int i = 0; i < 16; i++) sa += above[i], sl += left[i];
            if (x > 0 && yy > 0) dc = (sa + sl + 16) >> 5;
            else if (yy > 0) dc = (sa + 8) >> 4;
            else if (x > 0) dc = (sl + 8) >> 4;
            else dc = 128;
            int16_t res[256];
            int32_t coeff[256];
            for (int r = 0; r < 16; r++)
                for (int c = 0; c < 16; c++) res[r * 16 + c] = (int16_t)(y[(size_t)(yy + r) * stride + x + c] - dc);
            orc_fwd_txfm2d(res, coeff, 16, 0 /* DCT_DCT */, 2 /* TX_16X16 */, 8);
            int64_t s = 0;
            for (int i = 0; i < 256; i++) s += coeff[i] < 0 ? -(int64_t)coeff[i] : coeff[i];
            cost[my * mbw + mx] = s;
        }
}
