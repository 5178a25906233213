/* oracle.h — prototypes of the CPU restatement (TEST INFRASTRUCTURE; see the header of each .c file).
 * Built into oracle/liboracle.so by oracle/Makefile. Never linked into libsvtav1_b200.so. */
#ifndef SVT_B200_ORACLE_H
#define SVT_B200_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#include "../include/svt_av1_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
#define ORC_API __attribute__((visibility("default")))

/* ---- me_oracle.c ---- */
ORC_API uint32_t orc_nxm_sad(const uint8_t *src, uint32_t src_stride, const uint8_t *ref, uint32_t ref_stride,
                             uint32_t height, uint32_t width);
ORC_API void orc_sad_loop_kernel(const uint8_t *src, uint32_t src_stride, const uint8_t *ref,
                                 uint32_t ref_stride, uint32_t block_height, uint32_t block_width,
                                 uint64_t *best_sad, int16_t *x_search_center, int16_t *y_search_center,
                                 uint32_t src_stride_raw, int16_t search_area_width, int16_t search_area_height);
ORC_API void orc_ext_sad_calculation_8x8_16x16(const uint8_t *src, uint32_t src_stride, const uint8_t *ref,
                                               uint32_t ref_stride, uint32_t *p_best_sad_8x8,
                                               uint32_t *p_best_sad_16x16, uint32_t *p_best_mv8x8,
                                               uint32_t *p_best_mv16x16, uint32_t mv, uint32_t *p_sad16x16,
                                               uint32_t *p_sad8x8, uint8_t sub_sad);
ORC_API void orc_ext_sad_calculation_32x32_64x64(const uint32_t *p_sad16x16, uint32_t *p_best_sad_32x32,
                                                 uint32_t *p_best_sad_64x64, uint32_t *p_best_mv32x32,
                                                 uint32_t *p_best_mv64x64, uint32_t mv, uint32_t *p_sad32x32);
ORC_API void orc_ext_all_sad_calculation_8x8_16x16(const uint8_t *src, uint32_t src_stride, const uint8_t *ref,
                                                   uint32_t ref_stride, uint32_t mv, uint32_t *p_best_sad_8x8,
                                                   uint32_t *p_best_sad_16x16, uint32_t *p_best_mv8x8,
                                                   uint32_t *p_best_mv16x16, uint32_t p_eight_sad16x16[16][8],
                                                   uint32_t p_eight_sad8x8[64][8], uint8_t sub_sad);
ORC_API void orc_ext_eight_sad_calculation_32x32_64x64(uint32_t p_sad16x16[16][8], uint32_t *p_best_sad_32x32,
                                                       uint32_t *p_best_sad_64x64, uint32_t *p_best_mv32x32,
                                                       uint32_t *p_best_mv64x64, uint32_t mv,
                                                       uint32_t p_sad32x32[4][8]);
ORC_API void orc_me_sb(const SvtB200MeParams *p, const SvtB200MePlanes *src, const SvtB200MePlanes refs[2][4],
                       int sb_x, int sb_y, uint32_t *best_sad, uint32_t *best_mv, SvtB200HmeResult *hme_out,
                       int16_t *me_mv, uint8_t *me_cand, uint8_t *total_cand, uint32_t *rc_me_distortion);
ORC_API void orc_me_picture(const SvtB200MeParams *p, const SvtB200MePlanes *src,
                            const SvtB200MePlanes refs[2][4], uint32_t *best_sad, uint32_t *best_mv,
                            SvtB200HmeResult *hme, int16_t *me_mv, uint8_t *me_cand, uint8_t *total_cand,
                            uint32_t *rc_me_distortion);
ORC_API void orc_me_downsample(const SvtB200Plane *full, const SvtB200Plane *quarter, const SvtB200Plane *sixteenth,
                               const SvtB200MePlanes *planes, int filtered);
/* ---- cdef_oracle.c ---- */
ORC_API int32_t orc_cdef_find_dir(const uint16_t *img, int32_t stride, int32_t *var, int32_t coeff_shift);
ORC_API void orc_cdef_filter_block(uint8_t *dst8, uint16_t *dst16, int32_t dstride, const uint16_t *in,
                                   int32_t pri_strength, int32_t sec_strength, int32_t dir, int32_t pri_damping,
                                   int32_t sec_damping, int32_t bsize, int32_t coeff_shift);
ORC_API void orc_cdef_search(const SvtB200CdefSearchParams *p, const SvtB200Frame *recon,
                             const SvtB200Frame *source, const uint8_t *skip8, int32_t skip_stride, uint64_t *mse);
ORC_API void orc_cdef_apply(const SvtB200CdefApplyParams *p, const SvtB200Frame *recon, const SvtB200Frame *out,
                            const uint8_t *skip8, int32_t skip_stride, const int8_t *fb_strength_idx);
ORC_API int orc_cdef_strength_table(int pick_method, SvtB200CdefSearchParams *p);
ORC_API void orc_cdef_decide(const SvtB200CdefDecideParams *p, const uint64_t *mse, const uint8_t *skip8, int skip_stride,
                             SvtB200CdefDecision *out, int8_t *fb_strength_idx);
ORC_API int orc_cdef_decide_table(int pick_method, SvtB200CdefDecideParams *p);
/* ---- txfm_oracle.c ---- */
ORC_API const int32_t *orc_cospi_table(int bit);
ORC_API const int32_t *orc_sinpi_table(int bit);
ORC_API void orc_fdct(int32_t *x, int stride, int n, int cos_bit);
ORC_API void orc_idct(int32_t *x, int stride, int n, int cos_bit, int clamp_bit);
ORC_API void orc_fadst(int32_t *x, int stride, int n, int cos_bit);
ORC_API void orc_iadst(int32_t *x, int stride, int n, int cos_bit, int clamp_bit);
ORC_API void orc_fidentity(int32_t *x, int stride, int n);
ORC_API void orc_fwd_txfm2d(const int16_t *input, int32_t *output, uint32_t stride, int tx_type, int tx_size, int bit_depth);
ORC_API void orc_fwd_txfm2d_pf(const int16_t *input, int32_t *output, uint32_t stride, int tx_type, int tx_size,
                               int bit_depth, int shift);
ORC_API uint64_t orc_handle_transform64(int32_t *output, int tx_size);
ORC_API uint64_t orc_estimate_transform(const int16_t *res, uint32_t stride, int32_t *coeff, int tx_size, int bit_depth,
                                        int tx_type, int shape);
ORC_API void orc_inv_txfm2d_add(const int32_t *input, const uint16_t *pred, int32_t stride_r, uint16_t *recon,
                                int32_t stride_w, int tx_type, int tx_size, int bd);
ORC_API void orc_residual(const void *src, uint32_t src_stride, const void *pred, uint32_t pred_stride, int16_t *res,
                          uint32_t res_stride, uint32_t w, uint32_t h, int hbd);
ORC_API void orc_quantize_b(const int32_t *coeff, intptr_t n, const int16_t *zbin, const int16_t *round,
                            const int16_t *quant, const int16_t *quant_shift, int32_t *qcoeff, int32_t *dqcoeff,
                            const int16_t *dequant, uint16_t *eob_ptr, const int16_t *scan, const uint8_t *qm,
                            const uint8_t *iqm, int log_scale, int hbd);
ORC_API void orc_quantize_fp(const int32_t *coeff, intptr_t n, const int16_t *round, const int16_t *quant,
                             int32_t *qcoeff, int32_t *dqcoeff, const int16_t *dequant, uint16_t *eob_ptr,
                             const int16_t *scan, int log_scale, int hbd);
/* ---- dlf_oracle.c ---- */
ORC_API void orc_lpf_edge(void *s, int hbd, int across, int along, int len, int blimit, int limit, int thresh, int bd);
ORC_API void orc_dlf_frame(const SvtB200DlfParams *p, const SvtB200Frame *f, const SvtB200DlfMi *mi);
ORC_API void orc_frame_sse(const SvtB200Frame *a, const SvtB200Frame *b, uint64_t *sse);
ORC_API void orc_lf_level_lut(const SvtB200LfFrameInit *init, const int32_t levels[4], uint8_t lut[3][2][128]);
ORC_API void orc_pick_filter_level(const SvtB200LpfPickParams *p, const SvtB200Frame *recon, const SvtB200Frame *source,
                                   const SvtB200Frame *temp, const SvtB200DlfMi *mi, int32_t *levels_out);
/* ---- lr_oracle.c ---- */
ORC_API void orc_selfguided_restoration(const void *dgd, int hbd, int width, int height, int dgd_stride, int32_t *flt0,
                                        int32_t *flt1, int flt_stride, int sgr_params_idx, int bit_depth);
ORC_API void orc_apply_selfguided_restoration(const void *dat, int hbd, int width, int height, int stride, int eps,
                                              const int32_t *xqd, void *dst, int dst_stride, int bit_depth);
ORC_API void orc_wiener_convolve_add_src(const void *src, int hbd, ptrdiff_t src_stride, void *dst, ptrdiff_t dst_stride,
                                         const int16_t *filter_x, const int16_t *filter_y, int w, int h, int round_0,
                                         int round_1, int bd);
ORC_API void orc_lr_frame(const SvtB200LrFrameParams *p, const SvtB200Frame *cdef, const SvtB200Frame *dblk, const SvtB200Frame *out,
                          const SvtB200LrUnit *const units[3]);
/* ---- interp_oracle.c ---- */
ORC_API void orc_interp_kernel(int filter, int w, int subpel, int16_t out[8]);
ORC_API void orc_convolve(const void *src, int hbd, int src_stride, void *dst, int dst_stride, int w, int h, const int16_t *fx,
                          const int16_t *fy, int round_0, int round_1, int bd, int compound, int do_average, int use_jnt,
                          int fwd_offset, int bck_offset, uint16_t *conv_dst, int conv_stride);
ORC_API void orc_convolve8(const uint8_t *src, ptrdiff_t src_stride, uint8_t *dst, ptrdiff_t dst_stride, const int16_t *table,
                           int q0, int step, int w, int h, int vert);
ORC_API void orc_inter_predict(const SvtB200Frame *refs, int n_ref_frames, const SvtB200Frame *pred, const SvtB200InterJob *jobs,
                               int n_jobs);
/* ---- subpel_oracle.c ---- */
ORC_API void orc_upsampled_pred(uint8_t *comp_pred, int width, int height, int subpel_x_q3, int subpel_y_q3, const uint8_t *ref,
                                int ref_stride, int subpel_search);
ORC_API void orc_subpel_search(const SvtB200SubpelParams *p, const int32_t *mvcost0, const int32_t *mvcost1, const SvtB200Frame *src,
                               const SvtB200Frame *refs, int n_ref_frames, const SvtB200SubpelJob *jobs, int n_jobs,
                               SvtB200SubpelResult *results);
#ifdef __cplusplus
}
#endif
/* ---- temporal filter, planewise weighting (tf_oracle.c; EbTemporalFiltering.c:525-811, 829-1017, 1943-2050) ---- */
ORC_API const uint64_t *orc_exp2f_table(void);
ORC_API float orc_expf(float x);
ORC_API long orc_expf_mismatches(uint32_t lo_bits, uint32_t hi_bits, long *weight_mismatches);
ORC_API uint64_t orc_expf_checksum(uint32_t lo_bits, uint32_t hi_bits);
ORC_API int orc_tf_weight(uint64_t sum, int n, double block_error, double d_factor, double den);
ORC_API void orc_tf_den(int decay_control, const double *noise_levels, double den[3]);
ORC_API void orc_tf_block_factors(int split, const uint64_t err16[4], uint64_t err32, const int16_t mvx16[4], const int16_t mvy16[4],
                                  int16_t mvx32, int16_t mvy32, int min_frame_size, int hbd, double block_error[4], double d_factor[4]);
ORC_API void orc_tf_planewise(const void *y_src, int y_src_stride, const void *y_pre, int y_pre_stride, const void *u_src,
                              const void *v_src, int uv_src_stride, const void *u_pre, const void *v_pre, int uv_pre_stride,
                              unsigned bw, unsigned bh, int ss_x, int ss_y, const double den[3], const double block_error[4],
                              const double d_factor[4], int chroma, int bit_depth, uint32_t *y_accum, uint16_t *y_count,
                              uint32_t *u_accum, uint16_t *u_count, uint32_t *v_accum, uint16_t *v_count);
ORC_API void orc_tf_central(const void *pre, int pre_stride, unsigned w, unsigned h, int hbd, uint32_t *accum, uint16_t *count,
                            int acc_stride);
ORC_API uint64_t orc_tf_normalize(void *dst, int dst_stride, unsigned w, unsigned h, int hbd, const uint32_t *accum,
                                  const uint16_t *count, int acc_stride);

/* ---- picture-analysis block statistics (pa_oracle.c; EbPictureAnalysisProcess.c:287-378, 432-1003, 1005-2575, 2929-2974) ---- */
ORC_API void orc_sb_mean_variance(const uint8_t *y, int stride, int full_precision, uint8_t mean_out[85], uint16_t var_out[85]);
ORC_API void orc_sb_chroma_mean(const uint8_t *c, int stride, int full_precision, uint8_t mean_out[21]);
ORC_API uint16_t orc_picture_mean_variance(const uint8_t *y, int stride_y, const uint8_t *cb, const uint8_t *cr, int stride_c, int width,
                                           int height, int full_precision, uint8_t *y_mean, uint16_t *variance, uint8_t *cb_mean,
                                           uint8_t *cr_mean);

ORC_API void orc_ois_dc_picture(const uint8_t *y, int stride, int width, int height, int64_t *cost);

#endif
