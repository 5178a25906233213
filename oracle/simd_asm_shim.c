/* simd_asm_shim.c - BASELINE INFRASTRUCTURE (not product code, never linked into libsvtav1_b200.so).
 *
 * The reference's SIMD build needs nasm/yasm for 13 .asm files (CMakeLists.txt:48-69, enable_language(ASM_NASM));
 * this image has neither.  oracle/Makefile.simd therefore builds every *intrinsic* .c file of the reference
 * (ASM_SSE2 ... ASM_AVX512, compiled where they lie under /root/reference) and this file supplies the 36 symbols
 * that only the .asm files define, as plain-C code of our own with the semantics the callers expect:
 *   - intra predictors (intrapred_sse2.asm, highbd_intrapred_sse2_.asm)  -> the reference's own *_c predictors
 *   - svt_aom_subtract_block_sse2 (subtract_sse2.asm)                    -> svt_aom_subtract_block_c
 *   - picture_copy_kernel_sse2 (EbPictureOperators_SSE2.asm)             -> row memcpy
 *   - svt_aom_filter_block1d{4,8,16}_{v,h}2_ssse3, ..1d4_v8_sse2 (aom_subpixel_bilinear_ssse3.asm /
 *     aom_subpixel_8t_sse2.asm): 2-tap / 8-tap 1-D filters, round 64 >> 7, clip to 8 bits
 *   - svt_aom_highbd_calc{4x4,8x8,16x16}var_sse2 (highbd_variance_impl_sse2.asm): sum and SSE of a block
 *   - Log2f_ASM, RunEmms (x64RegisterUtil.asm)
 * None of these is one of the hot-path kernels this repository accelerates (SAD search, transforms, quantisers,
 * deblocking, CDEF, restoration all have intrinsic .c implementations and ARE the reference's AVX2/AVX-512 code);
 * the deviation "AVX2-minus-asm" is stated next to every CPU number that uses this build.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define PRED(name, T, extra_decl, extra_arg)                                                             \
    void name##_c(T *dst, ptrdiff_t stride, const T *above, const T *left extra_decl);                   \
    void name##_sse2(T *dst, ptrdiff_t stride, const T *above, const T *left extra_decl) {               \
        name##_c(dst, stride, above, left extra_arg);                                                    \
    }
#define COMMA_BD_DECL , int32_t bd
#define COMMA_BD_ARG , bd
#define LBD3(kind) \
    PRED(svt_aom_##kind##_predictor_4x4, uint8_t, , ) PRED(svt_aom_##kind##_predictor_8x8, uint8_t, , ) \
    PRED(svt_aom_##kind##_predictor_16x16, uint8_t, , )
LBD3(dc) LBD3(dc_128) LBD3(dc_left) LBD3(dc_top) LBD3(v) LBD3(h)
PRED(svt_aom_highbd_dc_predictor_4x4, uint16_t, COMMA_BD_DECL, COMMA_BD_ARG)
PRED(svt_aom_highbd_dc_predictor_8x8, uint16_t, COMMA_BD_DECL, COMMA_BD_ARG)
PRED(svt_aom_highbd_v_predictor_4x4, uint16_t, COMMA_BD_DECL, COMMA_BD_ARG)
PRED(svt_aom_highbd_v_predictor_8x8, uint16_t, COMMA_BD_DECL, COMMA_BD_ARG)

void svt_aom_subtract_block_c(int rows, int cols, int16_t *diff_ptr, ptrdiff_t diff_stride, const uint8_t *src_ptr,
                              ptrdiff_t src_stride, const uint8_t *pred_ptr, ptrdiff_t pred_stride);
void svt_aom_subtract_block_sse2(int rows, int cols, int16_t *diff_ptr, ptrdiff_t diff_stride, const uint8_t *src_ptr,
                                 ptrdiff_t src_stride, const uint8_t *pred_ptr, ptrdiff_t pred_stride) {
    svt_aom_subtract_block_c(rows, cols, diff_ptr, diff_stride, src_ptr, src_stride, pred_ptr, pred_stride);
}

void picture_copy_kernel_sse2(uint8_t *src, uint32_t src_stride, uint8_t *dst, uint32_t dst_stride,
                              uint32_t area_width, uint32_t area_height) {
    for (uint32_t y = 0; y < area_height; y++) memcpy(dst + (size_t)y * dst_stride, src + (size_t)y * src_stride, area_width);
}

static inline uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }

/* 2-tap (bilinear) 1-D filters: taps filter[3], filter[4]; src points AT the first sample. */
#define BILINEAR(W, dir, step)                                                                            \
    void svt_aom_filter_block1d##W##_##dir##2_ssse3(const uint8_t *src, ptrdiff_t src_pitch, uint8_t *out,\
                                                    ptrdiff_t out_pitch, uint32_t h, const int16_t *f) {  \
        for (uint32_t y = 0; y < h; y++)                                                                   \
            for (int x = 0; x < W; x++) {                                                                  \
                const uint8_t *s = src + (ptrdiff_t)y * src_pitch + x;                                     \
                out[(ptrdiff_t)y * out_pitch + x] = clip8((s[0] * f[3] + s[step] * f[4] + 64) >> 7);       \
            }                                                                                              \
    }
BILINEAR(4, v, src_pitch) BILINEAR(8, v, src_pitch) BILINEAR(16, v, src_pitch)
BILINEAR(4, h, 1) BILINEAR(8, h, 1) BILINEAR(16, h, 1)

/* 8-tap vertical, 4 wide; src points 3 rows ABOVE the first output row (the caller passes src - 3*stride). */
void svt_aom_filter_block1d4_v8_sse2(const uint8_t *src, ptrdiff_t src_pitch, uint8_t *out, ptrdiff_t out_pitch,
                                     uint32_t h, const int16_t *f) {
    for (uint32_t y = 0; y < h; y++)
        for (int x = 0; x < 4; x++) {
            int sum = 64;
            for (int k = 0; k < 8; k++) sum += src[((ptrdiff_t)y + k) * src_pitch + x] * f[k];
            out[(ptrdiff_t)y * out_pitch + x] = clip8(sum >> 7);
        }
}

#define HBD_VAR(N)                                                                                         \
    uint32_t svt_aom_highbd_calc##N##x##N##var_sse2(const uint16_t *src, int32_t src_stride,              \
                                                    const uint16_t *ref, int32_t ref_stride, uint32_t *sse, \
                                                    int32_t *sum) {                                        \
        int32_t s = 0; uint32_t q = 0;                                                                     \
        for (int y = 0; y < N; y++)                                                                        \
            for (int x = 0; x < N; x++) {                                                                  \
                const int d = (int)src[y * src_stride + x] - (int)ref[y * ref_stride + x];                 \
                s += d; q += (uint32_t)(d * d);                                                            \
            }                                                                                              \
        *sum = s; *sse = q; return 0;                                                                      \
    }
HBD_VAR(4) HBD_VAR(8) HBD_VAR(16)

uint32_t Log2f_ASM(uint32_t x) { return x ? 31u - (uint32_t)__builtin_clz(x) : 0u; }
void     RunEmms(void) {}
