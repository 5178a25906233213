/*
 * rtcd_install.c — the reference-side binding of INTEGRATION.md section 1, as real code: every *_cuda drop-in of
 * include/svt_av1_b200.h assigned to the RTCD function pointer of the UNMODIFIED reference it replaces.
 * TEST INFRASTRUCTURE (built by Makefile.ref into oracle/_ref/librefcuda.so against the reference's own headers and
 * library): tests/test_rtcd_install_gpu.py runs the reference's C loops (deblocking frame, CDEF search / apply, the EncDec
 * chain, restoration, inter prediction, sub-pel search) with these pointers installed and requires the outputs to be
 * identical to the same loops on the C pointers.
 *
 * Compiling this file is itself a check: EXACT() initialises a variable of the pointer's own type, so the 199 drop-ins
 * listed there have signatures IDENTICAL to the reference's (gcc -Werror=incompatible-pointer-types).  The 26 under
 * CAST() differ only in how a pointed-to struct or an int-sized enum is spelled in our header (mirror structs
 * SvtB200InterpFilterParams / SvtB200ConvolveParams, `const void *` for CdefList / SgrParamsType / TxfmParam /
 * MacroBlockD, int32_t for AomBitDepth): same size, same registers.
 */
#include "EbDefinitions.h"
#include "aom_dsp_rtcd.h"
#include "common_dsp_rtcd.h"

#include "../include/svt_av1_b200.h"

/* svt_av1_inv_txfm_add_cuda reads the reference's TxfmParam through a layout view (txfm.cu TxfmParamView: packed 1-byte
 * TxType / TxSize enums, then int32 lossless, bd, is_hbd, a 1-byte TxSetType, int32 eob).  This file sees both headers:
 * pin the layout at compile time so a build of the reference with other enum packing cannot be misread silently. */
#include <stddef.h>
_Static_assert(sizeof(((TxfmParam *)0)->tx_type) == 1 && sizeof(((TxfmParam *)0)->tx_size) == 1, "TxType / TxSize must be packed 1-byte enums");
_Static_assert(offsetof(TxfmParam, tx_type) == 0 && offsetof(TxfmParam, tx_size) == 1, "TxfmParam: tx_type, tx_size first");
_Static_assert(offsetof(TxfmParam, lossless) == 4 && offsetof(TxfmParam, bd) == 8 && offsetof(TxfmParam, is_hbd) == 12,
               "TxfmParam: lossless / bd / is_hbd at 4 / 8 / 12");
_Static_assert(offsetof(TxfmParam, eob) == 20 && sizeof(TxfmParam) == 24, "TxfmParam: eob at 20, 24 bytes");

#include "../integration/svt_cuda_tf_shim.h"

#define REFH_API __attribute__((visibility("default")))
typedef void (*AnyFn)(void);
#define MAX_SLOTS 256
static AnyFn *g_slot[MAX_SLOTS];
static AnyFn g_saved[MAX_SLOTS];
static int g_n = 0, g_installed = 0;

#define EXACT(name)                                         \
    do {                                                    \
        __typeof__(name) fn__ = name##_cuda;                \
        g_slot[g_n] = (AnyFn *)&name;                       \
        g_saved[g_n++] = (AnyFn)name;                       \
        name = fn__;                                        \
    } while (0);
#define CAST(name)                                          \
    do {                                                    \
        g_slot[g_n] = (AnyFn *)&name;                       \
        g_saved[g_n++] = (AnyFn)name;                       \
        name = (__typeof__(name))(AnyFn)name##_cuda;        \
    } while (0);

/* call after setup_common_rtcd_internal() / setup_rtcd_internal() (svt_av1_enc_init, EbEncHandle.c:1144-1145) */
REFH_API int svt_cuda_install_rtcd(void) {
    if (g_installed) return g_n;
    if (svt_b200_device_count() <= 0) return -1; /* no CPU fallback: the caller aborts */
    g_n = 0;
    EXACT(sad_16b_kernel) EXACT(svt_aom_convolve8_horiz) EXACT(svt_aom_convolve8_vert)
    EXACT(svt_aom_highbd_lpf_horizontal_14) EXACT(svt_aom_highbd_lpf_horizontal_4)
    EXACT(svt_aom_highbd_lpf_horizontal_6) EXACT(svt_aom_highbd_lpf_horizontal_8)
    EXACT(svt_aom_highbd_lpf_vertical_14) EXACT(svt_aom_highbd_lpf_vertical_4) EXACT(svt_aom_highbd_lpf_vertical_6)
    EXACT(svt_aom_highbd_lpf_vertical_8) EXACT(svt_aom_highbd_quantize_b) EXACT(svt_aom_highbd_subtract_block)
    EXACT(svt_aom_lpf_horizontal_14) EXACT(svt_aom_lpf_horizontal_4) EXACT(svt_aom_lpf_horizontal_6)
    EXACT(svt_aom_lpf_horizontal_8) EXACT(svt_aom_lpf_vertical_14) EXACT(svt_aom_lpf_vertical_4)
    EXACT(svt_aom_lpf_vertical_6) EXACT(svt_aom_lpf_vertical_8) EXACT(svt_aom_mse16x16) EXACT(svt_aom_quantize_b)
    EXACT(svt_aom_sad128x128) EXACT(svt_aom_sad128x128x4d) EXACT(svt_aom_sad128x64) EXACT(svt_aom_sad128x64x4d)
    EXACT(svt_aom_sad16x16) EXACT(svt_aom_sad16x16x4d) EXACT(svt_aom_sad16x32) EXACT(svt_aom_sad16x32x4d)
    EXACT(svt_aom_sad16x4) EXACT(svt_aom_sad16x4x4d) EXACT(svt_aom_sad16x64) EXACT(svt_aom_sad16x64x4d)
    EXACT(svt_aom_sad16x8) EXACT(svt_aom_sad16x8x4d) EXACT(svt_aom_sad32x16) EXACT(svt_aom_sad32x16x4d)
    EXACT(svt_aom_sad32x32) EXACT(svt_aom_sad32x32x4d) EXACT(svt_aom_sad32x64) EXACT(svt_aom_sad32x64x4d)
    EXACT(svt_aom_sad32x8) EXACT(svt_aom_sad32x8x4d) EXACT(svt_aom_sad4x16) EXACT(svt_aom_sad4x16x4d)
    EXACT(svt_aom_sad4x4) EXACT(svt_aom_sad4x4x4d) EXACT(svt_aom_sad4x8) EXACT(svt_aom_sad4x8x4d)
    EXACT(svt_aom_sad64x128) EXACT(svt_aom_sad64x128x4d) EXACT(svt_aom_sad64x16) EXACT(svt_aom_sad64x16x4d)
    EXACT(svt_aom_sad64x32) EXACT(svt_aom_sad64x32x4d) EXACT(svt_aom_sad64x64) EXACT(svt_aom_sad64x64x4d)
    EXACT(svt_aom_sad8x16) EXACT(svt_aom_sad8x16x4d) EXACT(svt_aom_sad8x32) EXACT(svt_aom_sad8x32x4d)
    EXACT(svt_aom_sad8x4) EXACT(svt_aom_sad8x4x4d) EXACT(svt_aom_sad8x8) EXACT(svt_aom_sad8x8x4d) EXACT(svt_aom_satd)
    EXACT(svt_aom_subtract_block) EXACT(svt_aom_variance128x128) EXACT(svt_aom_variance128x64)
    EXACT(svt_aom_variance16x16) EXACT(svt_aom_variance16x32) EXACT(svt_aom_variance16x4) EXACT(svt_aom_variance16x64)
    EXACT(svt_aom_variance16x8) EXACT(svt_aom_variance32x16) EXACT(svt_aom_variance32x32) EXACT(svt_aom_variance32x64)
    EXACT(svt_aom_variance32x8) EXACT(svt_aom_variance4x16) EXACT(svt_aom_variance4x4) EXACT(svt_aom_variance4x8)
    EXACT(svt_aom_variance64x128) EXACT(svt_aom_variance64x16) EXACT(svt_aom_variance64x32)
    EXACT(svt_aom_variance64x64) EXACT(svt_aom_variance8x16) EXACT(svt_aom_variance8x32) EXACT(svt_aom_variance8x4)
    EXACT(svt_aom_variance8x8) EXACT(svt_apply_selfguided_restoration) EXACT(svt_av1_block_error)
    EXACT(svt_av1_compute_stats) EXACT(svt_av1_fwd_txfm2d_16x16) EXACT(svt_av1_fwd_txfm2d_16x16_N2)
    EXACT(svt_av1_fwd_txfm2d_16x16_N4) EXACT(svt_av1_fwd_txfm2d_16x32) EXACT(svt_av1_fwd_txfm2d_16x32_N2)
    EXACT(svt_av1_fwd_txfm2d_16x32_N4) EXACT(svt_av1_fwd_txfm2d_16x4) EXACT(svt_av1_fwd_txfm2d_16x4_N2)
    EXACT(svt_av1_fwd_txfm2d_16x4_N4) EXACT(svt_av1_fwd_txfm2d_16x64) EXACT(svt_av1_fwd_txfm2d_16x64_N2)
    EXACT(svt_av1_fwd_txfm2d_16x64_N4) EXACT(svt_av1_fwd_txfm2d_16x8) EXACT(svt_av1_fwd_txfm2d_16x8_N2)
    EXACT(svt_av1_fwd_txfm2d_16x8_N4) EXACT(svt_av1_fwd_txfm2d_32x16) EXACT(svt_av1_fwd_txfm2d_32x16_N2)
    EXACT(svt_av1_fwd_txfm2d_32x16_N4) EXACT(svt_av1_fwd_txfm2d_32x32) EXACT(svt_av1_fwd_txfm2d_32x32_N2)
    EXACT(svt_av1_fwd_txfm2d_32x32_N4) EXACT(svt_av1_fwd_txfm2d_32x64) EXACT(svt_av1_fwd_txfm2d_32x64_N2)
    EXACT(svt_av1_fwd_txfm2d_32x64_N4) EXACT(svt_av1_fwd_txfm2d_32x8) EXACT(svt_av1_fwd_txfm2d_32x8_N2)
    EXACT(svt_av1_fwd_txfm2d_32x8_N4) EXACT(svt_av1_fwd_txfm2d_4x16) EXACT(svt_av1_fwd_txfm2d_4x16_N2)
    EXACT(svt_av1_fwd_txfm2d_4x16_N4) EXACT(svt_av1_fwd_txfm2d_4x4) EXACT(svt_av1_fwd_txfm2d_4x4_N2)
    EXACT(svt_av1_fwd_txfm2d_4x4_N4) EXACT(svt_av1_fwd_txfm2d_4x8) EXACT(svt_av1_fwd_txfm2d_4x8_N2)
    EXACT(svt_av1_fwd_txfm2d_4x8_N4) EXACT(svt_av1_fwd_txfm2d_64x16) EXACT(svt_av1_fwd_txfm2d_64x16_N2)
    EXACT(svt_av1_fwd_txfm2d_64x16_N4) EXACT(svt_av1_fwd_txfm2d_64x32) EXACT(svt_av1_fwd_txfm2d_64x32_N2)
    EXACT(svt_av1_fwd_txfm2d_64x32_N4) EXACT(svt_av1_fwd_txfm2d_64x64) EXACT(svt_av1_fwd_txfm2d_64x64_N2)
    EXACT(svt_av1_fwd_txfm2d_64x64_N4) EXACT(svt_av1_fwd_txfm2d_8x16) EXACT(svt_av1_fwd_txfm2d_8x16_N2)
    EXACT(svt_av1_fwd_txfm2d_8x16_N4) EXACT(svt_av1_fwd_txfm2d_8x32) EXACT(svt_av1_fwd_txfm2d_8x32_N2)
    EXACT(svt_av1_fwd_txfm2d_8x32_N4) EXACT(svt_av1_fwd_txfm2d_8x4) EXACT(svt_av1_fwd_txfm2d_8x4_N2)
    EXACT(svt_av1_fwd_txfm2d_8x4_N4) EXACT(svt_av1_fwd_txfm2d_8x8) EXACT(svt_av1_fwd_txfm2d_8x8_N2)
    EXACT(svt_av1_fwd_txfm2d_8x8_N4) EXACT(svt_av1_highbd_quantize_fp) EXACT(svt_av1_inv_txfm2d_add_16x16)
    EXACT(svt_av1_inv_txfm2d_add_16x32) EXACT(svt_av1_inv_txfm2d_add_16x4) EXACT(svt_av1_inv_txfm2d_add_16x64)
    EXACT(svt_av1_inv_txfm2d_add_16x8) EXACT(svt_av1_inv_txfm2d_add_32x16) EXACT(svt_av1_inv_txfm2d_add_32x32)
    EXACT(svt_av1_inv_txfm2d_add_32x64) EXACT(svt_av1_inv_txfm2d_add_32x8) EXACT(svt_av1_inv_txfm2d_add_4x16)
    EXACT(svt_av1_inv_txfm2d_add_4x4) EXACT(svt_av1_inv_txfm2d_add_4x8) EXACT(svt_av1_inv_txfm2d_add_64x16)
    EXACT(svt_av1_inv_txfm2d_add_64x32) EXACT(svt_av1_inv_txfm2d_add_64x64) EXACT(svt_av1_inv_txfm2d_add_8x16)
    EXACT(svt_av1_inv_txfm2d_add_8x32) EXACT(svt_av1_inv_txfm2d_add_8x4) EXACT(svt_av1_inv_txfm2d_add_8x8)
    EXACT(svt_av1_quantize_fp) EXACT(svt_av1_quantize_fp_32x32) EXACT(svt_av1_quantize_fp_64x64)
    EXACT(svt_av1_selfguided_restoration) EXACT(svt_cdef_filter_block) EXACT(svt_cdef_find_dir)
    EXACT(svt_copy_rect8_8bit_to_16bit) EXACT(svt_ext_all_sad_calculation_8x8_16x16)
    EXACT(svt_ext_eight_sad_calculation_32x32_64x64) EXACT(svt_ext_sad_calculation_32x32_64x64)
    EXACT(svt_ext_sad_calculation_8x8_16x16) EXACT(svt_full_distortion_kernel16_bits)
    EXACT(svt_full_distortion_kernel32_bits) EXACT(svt_full_distortion_kernel_cbf_zero32_bits)
    EXACT(svt_handle_transform16x64) EXACT(svt_handle_transform32x64) EXACT(svt_handle_transform64x16)
    EXACT(svt_handle_transform64x32) EXACT(svt_handle_transform64x64) EXACT(svt_initialize_buffer_32bits)
    EXACT(svt_nxm_sad_kernel) EXACT(svt_nxm_sad_kernel_sub_sampled) EXACT(svt_residual_kernel16bit)
    EXACT(svt_residual_kernel8bit) EXACT(svt_sad_loop_kernel) EXACT(svt_spatial_full_distortion_kernel)
    CAST(svt_aom_upsampled_pred) CAST(svt_av1_compute_stats_highbd) CAST(svt_av1_convolve_2d_copy_sr)
    CAST(svt_av1_convolve_2d_sr) CAST(svt_av1_convolve_x_sr) CAST(svt_av1_convolve_y_sr)
    CAST(svt_av1_highbd_convolve_2d_copy_sr) CAST(svt_av1_highbd_convolve_2d_sr) CAST(svt_av1_highbd_convolve_x_sr)
    CAST(svt_av1_highbd_convolve_y_sr) CAST(svt_av1_highbd_jnt_convolve_2d) CAST(svt_av1_highbd_jnt_convolve_2d_copy)
    CAST(svt_av1_highbd_jnt_convolve_x) CAST(svt_av1_highbd_jnt_convolve_y) CAST(svt_av1_highbd_pixel_proj_error)
    CAST(svt_av1_highbd_wiener_convolve_add_src) CAST(svt_av1_inv_txfm_add) CAST(svt_av1_jnt_convolve_2d)
    CAST(svt_av1_jnt_convolve_2d_copy) CAST(svt_av1_jnt_convolve_x) CAST(svt_av1_jnt_convolve_y)
    CAST(svt_av1_lowbd_pixel_proj_error) CAST(svt_av1_wiener_convolve_add_src) CAST(svt_compute_cdef_dist_16bit)
    CAST(svt_compute_cdef_dist_8bit) CAST(svt_get_proj_subspace)
    EXACT(svt_av1_apply_temporal_filter_planewise) EXACT(svt_av1_apply_temporal_filter_planewise_hbd)
    g_installed = 1;
    return g_n;
}

/* back to whatever was installed before (the tests compare both) */
REFH_API void svt_cuda_uninstall_rtcd(void) {
    if (!g_installed) return;
    for (int i = 0; i < g_n; i++) *g_slot[i] = g_saved[i];
    g_installed = 0;
}
