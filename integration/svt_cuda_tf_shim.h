/* svt_cuda_tf_shim.h - the MeContext-reading drop-ins for the reference's two temporal-filter RTCD pointers
 * (svt_av1_apply_temporal_filter_planewise / _hbd, aom_dsp_rtcd.c:365-366).  Compiled next to the reference's headers:
 * included by integration/svt_cuda_backend.c (SVT_CUDA_TF=1) and by oracle/rtcd_install.c (the test installer). */
#ifndef SVT_CUDA_TF_SHIM_H
#define SVT_CUDA_TF_SHIM_H
/* The two temporal-filter pointers take the reference's MeContext: the shim below reads the fields the C functions read
 * (EbTemporalFiltering.c:696-734), computes the block-level terms with the reference's own expressions on the host (they
 * are per 16x16 quadrant / per plane: sqrtf, powf, log1p of a handful of numbers) and hands plain numbers to the library. */
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include "EbMotionEstimationContext.h"
#include "EbTemporalFiltering.h"
static void tf_shim(struct MeContext *ctx, int bit_depth, const void *y_src, int y_src_stride, const void *y_pre, int y_pre_stride,
                    const void *u_src, const void *v_src, int uv_src_stride, const void *u_pre, const void *v_pre, int uv_pre_stride,
                    unsigned int block_width, unsigned int block_height, int ss_x, int ss_y, const double *noise_levels,
                    const int decay_control, uint32_t *y_accum, uint16_t *y_count, uint32_t *u_accum, uint16_t *u_count,
                    uint32_t *v_accum, uint16_t *v_count) {
    if (ss_x != 1 || ss_y != 1) {
        fprintf(stderr, "svt_av1_apply_temporal_filter_planewise_cuda: 4:2:0 only\n");
        abort();
    }
#ifdef SVT_CUDA_TF_COUNT
    SVT_CUDA_TF_COUNT();
#endif
    double den[3], block_error[4], d_factor[4];
    for (int p = 0; p < 3; p++) {
        double n_decay = (double)decay_control * (0.7 + log1p(noise_levels[p]));
        den[p] = 2 * n_decay * n_decay;
    }
    const int idx_32x32 = ctx->tf_block_col + ctx->tf_block_row * 2, hbd = bit_depth > 8;
    for (int q = 0; q < 4; q++) {
        MV mv;
        if (ctx->tf_32x32_block_split_flag[idx_32x32]) {
            const uint64_t e = ctx->tf_16x16_block_error[idx_32x32 * 4 + q];
            block_error[q] = (double)(hbd ? e >> 4 : e) / 256;
            mv.col = ctx->tf_16x16_mv_x[idx_32x32 * 4 + q];
            mv.row = ctx->tf_16x16_mv_y[idx_32x32 * 4 + q];
        } else {
            const uint64_t e = ctx->tf_32x32_block_error[idx_32x32];
            block_error[q] = (double)(hbd ? e >> 4 : e) / 1024;
            mv.col = ctx->tf_32x32_mv_x[idx_32x32];
            mv.row = ctx->tf_32x32_mv_y[idx_32x32];
        }
        const float  distance           = sqrtf(powf(mv.row, 2) + powf(mv.col, 2));
        const double distance_threshold = (double)AOMMAX(ctx->min_frame_size * TF_SEARCH_DISTANCE_THRESHOLD, 1);
        d_factor[q]                     = AOMMAX(distance / distance_threshold, 1);
    }
    if (svt_b200_tf_planewise_block_host(bit_depth, ctx->tf_chroma, y_src, y_src_stride, y_pre, y_pre_stride, u_src, v_src, uv_src_stride,
                                         u_pre, v_pre, uv_pre_stride, block_width, block_height, den, block_error, d_factor, y_accum,
                                         y_count, u_accum, u_count, v_accum, v_count)) {
        fprintf(stderr, "svt_av1_apply_temporal_filter_planewise_cuda: %s\n", svt_b200_last_error());
        abort(); /* no CPU fallback */
    }
}
static void svt_av1_apply_temporal_filter_planewise_cuda(struct MeContext *context_ptr, const uint8_t *y_src, int y_src_stride,
                                                         const uint8_t *y_pre, int y_pre_stride, const uint8_t *u_src,
                                                         const uint8_t *v_src, int uv_src_stride, const uint8_t *u_pre,
                                                         const uint8_t *v_pre, int uv_pre_stride, unsigned int block_width,
                                                         unsigned int block_height, int ss_x, int ss_y, const double *noise_levels,
                                                         const int decay_control, uint32_t *y_accum, uint16_t *y_count,
                                                         uint32_t *u_accum, uint16_t *u_count, uint32_t *v_accum, uint16_t *v_count) {
    tf_shim(context_ptr, 8, y_src, y_src_stride, y_pre, y_pre_stride, u_src, v_src, uv_src_stride, u_pre, v_pre, uv_pre_stride, block_width,
            block_height, ss_x, ss_y, noise_levels, decay_control, y_accum, y_count, u_accum, u_count, v_accum, v_count);
}
static void svt_av1_apply_temporal_filter_planewise_hbd_cuda(struct MeContext *context_ptr, const uint16_t *y_src, int y_src_stride,
                                                             const uint16_t *y_pre, int y_pre_stride, const uint16_t *u_src,
                                                             const uint16_t *v_src, int uv_src_stride, const uint16_t *u_pre,
                                                             const uint16_t *v_pre, int uv_pre_stride, unsigned int block_width,
                                                             unsigned int block_height, int ss_x, int ss_y,
                                                             const double *noise_levels, const int decay_control, uint32_t *y_accum,
                                                             uint16_t *y_count, uint32_t *u_accum, uint16_t *u_count,
                                                             uint32_t *v_accum, uint16_t *v_count, uint32_t encoder_bit_depth) {
    tf_shim(context_ptr, (int)encoder_bit_depth, y_src, y_src_stride, y_pre, y_pre_stride, u_src, v_src, uv_src_stride, u_pre, v_pre,
            uv_pre_stride, block_width, block_height, ss_x, ss_y, noise_levels, decay_control, y_accum, y_count, u_accum, u_count, v_accum,
            v_count);
}

#endif
