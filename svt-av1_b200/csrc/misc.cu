// misc.cu — the small reductions that sit next to the transform chain and the single-position SAD family, as RTCD
// drop-ins (host pointers).  Replaces (reference files under Source/Lib):
//   svt_aom_subtract_block / svt_aom_highbd_subtract_block          Common/Codec/EbInterPrediction.c:65
//   svt_full_distortion_kernel32_bits / _cbf_zero32_bits            Common/Codec/EbPictureOperators.c:156-231
//   svt_spatial_full_distortion_kernel / svt_full_distortion_kernel16_bits  Common/C_DEFAULT/EbPictureOperators_C.c:65, EbPictureOperators.c:182
//   svt_aom_satd, svt_av1_block_error                               Common/Codec/common_dsp_rtcd.c:47-69
//   svt_aom_sadMxN, svt_aom_sadMxNx4d (22 sizes), svt_nxm_sad_kernel_sub_sampled, sad_16b_kernel
//                                                                   Encoder/C_DEFAULT/sad_av1.c, EbComputeSAD_C.c:39-56
// One block-wide reduction kernel serves all of them (mode selects the per-element term); these calls are tiny, so
// the wrappers exist for pointer-table completeness and parity tests, not for throughput.
#include "common.cuh"

using namespace svtb200;

namespace {

enum { M_SAD8 = 0, M_SAD16, M_SSE8, M_SSE16, M_DIST32, M_SQ32, M_SATD, M_BLKERR };

// a, b: packed w x h arrays (element size by mode). out[0], out[1]: 64-bit results.
__global__ void __launch_bounds__(256) reduce2_kernel(const void *a, const void *b, int n, int mode, unsigned long long *out) {
    unsigned long long r0 = 0, r1 = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        switch (mode) {
        case M_SAD8: r0 += (unsigned)abs((int)((const uint8_t *)a)[i] - (int)((const uint8_t *)b)[i]); break;
        case M_SAD16: r0 += (unsigned)abs((int)((const uint16_t *)a)[i] - (int)((const uint16_t *)b)[i]); break;
        case M_SSE8: {
            const long long d = (long long)((const uint8_t *)a)[i] - ((const uint8_t *)b)[i];
            r0 += (unsigned long long)(d * d);
            break;
        }
        case M_SSE16: {
            const long long d = (long long)((const uint16_t *)a)[i] - ((const uint16_t *)b)[i];
            r0 += (unsigned long long)(d * d);
            break;
        }
        case M_DIST32: {
            const long long c = ((const int32_t *)a)[i], d = c - (long long)((const int32_t *)b)[i];
            r0 += (unsigned long long)(d * d);
            r1 += (unsigned long long)(c * c);
            break;
        }
        case M_SQ32: {
            const long long c = ((const int32_t *)a)[i];
            r0 += (unsigned long long)(c * c);
            break;
        }
        case M_SATD: r0 += (unsigned long long)(long long)abs(((const int32_t *)a)[i]); break;
        default: { // svt_av1_block_error_c: int products (32-bit), 64-bit sums
            const int c = ((const int32_t *)a)[i], d = c - ((const int32_t *)b)[i];
            r0 += (unsigned long long)(long long)(int)((unsigned)d * (unsigned)d);
            r1 += (unsigned long long)(long long)(int)((unsigned)c * (unsigned)c);
            break;
        }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        r0 += __shfl_xor_sync(0xffffffffu, r0, o);
        r1 += __shfl_xor_sync(0xffffffffu, r1, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], r0);
        atomicAdd(&out[1], r1);
    }
}
template <typename T>
__global__ void subtract_kernel(const T *src, const T *pred, int16_t *diff, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) diff[i] = (int16_t)((int)src[i] - (int)pred[i]);
}

static void pack(uint8_t *dst, const void *src, size_t stride_elems, int w, int h, int esz) {
    for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * w * esz, (const uint8_t *)src + (size_t)y * stride_elems * esz, (size_t)w * esz);
}

// returns {r0, r1}
static void reduce2(const void *a, size_t sa, const void *b, size_t sb, int w, int h, int esz, int mode, uint64_t res[2]) {
    ThreadCtx &c = tls();
    const size_t nb = (size_t)w * h * esz, off_b = (nb + 15) & ~(size_t)15, off_o = 2 * off_b;
    c.reserve(off_o + 16);
    pack(c.h, a, sa, w, h, esz);
    if (b) pack(c.h + off_b, b, sb, w, h, esz);
    memset(c.h + off_o, 0, 16);
    SVTB_CUDA_FATAL(cudaMemcpyAsync(c.d, c.h, off_o + 16, cudaMemcpyHostToDevice, c.stream));
    SVTB_LAUNCH(reduce2_kernel, 1, 256, 0, c.stream, (const void *)c.d, (const void *)(c.d + off_b), w * h, mode,
                (unsigned long long *)(c.d + off_o));
    SVTB_CUDA_FATAL(cudaMemcpyAsync(c.h + off_o, c.d + off_o, 16, cudaMemcpyDeviceToHost, c.stream));
    SVTB_CUDA_FATAL(cudaStreamSynchronize(c.stream));
    memcpy(res, c.h + off_o, 16);
}

static void subtract(int rows, int cols, int16_t *diff, ptrdiff_t ds, const void *src, ptrdiff_t ss, const void *pred, ptrdiff_t ps,
                     int esz) {
    ThreadCtx &c = tls();
    const size_t nb = (size_t)rows * cols * esz, off_p = (nb + 15) & ~(size_t)15, off_d = 2 * off_p;
    c.reserve(off_d + (size_t)rows * cols * 2);
    pack(c.h, src, ss, cols, rows, esz);
    pack(c.h + off_p, pred, ps, cols, rows, esz);
    SVTB_CUDA_FATAL(cudaMemcpyAsync(c.d, c.h, off_d, cudaMemcpyHostToDevice, c.stream));
    if (esz == 2)
        SVTB_LAUNCH(subtract_kernel<uint16_t>, 4, 256, 0, c.stream, (const uint16_t *)c.d, (const uint16_t *)(c.d + off_p),
                    (int16_t *)(c.d + off_d), rows * cols);
    else
        SVTB_LAUNCH(subtract_kernel<uint8_t>, 4, 256, 0, c.stream, (const uint8_t *)c.d, (const uint8_t *)(c.d + off_p),
                    (int16_t *)(c.d + off_d), rows * cols);
    SVTB_CUDA_FATAL(cudaMemcpyAsync(c.h + off_d, c.d + off_d, (size_t)rows * cols * 2, cudaMemcpyDeviceToHost, c.stream));
    SVTB_CUDA_FATAL(cudaStreamSynchronize(c.stream));
    for (int y = 0; y < rows; y++) memcpy(diff + (ptrdiff_t)y * ds, c.h + off_d + (size_t)y * cols * 2, (size_t)cols * 2);
}

} // namespace

extern "C" {

void svt_aom_subtract_block_cuda(int rows, int cols, int16_t *diff_ptr, ptrdiff_t diff_stride, const uint8_t *src_ptr,
                                 ptrdiff_t src_stride, const uint8_t *pred_ptr, ptrdiff_t pred_stride) {
    subtract(rows, cols, diff_ptr, diff_stride, src_ptr, src_stride, pred_ptr, pred_stride, 1);
}
void svt_aom_highbd_subtract_block_cuda(int rows, int cols, int16_t *diff_ptr, ptrdiff_t diff_stride, const uint8_t *src_ptr,
                                        ptrdiff_t src_stride, const uint8_t *pred_ptr, ptrdiff_t pred_stride, int bd) {
    (void)bd; // CONVERT_TO_SHORTPTR convention of the reference
    subtract(rows, cols, diff_ptr, diff_stride, (const void *)(((uintptr_t)src_ptr) << 1), src_stride,
             (const void *)(((uintptr_t)pred_ptr) << 1), pred_stride, 2);
}
void svt_full_distortion_kernel32_bits_cuda(int32_t *coeff, uint32_t coeff_stride, int32_t *recon_coeff, uint32_t recon_coeff_stride,
                                            uint64_t distortion_result[2], uint32_t area_width, uint32_t area_height) {
    reduce2(coeff, coeff_stride, recon_coeff, recon_coeff_stride, area_width, area_height, 4, M_DIST32, distortion_result);
}
void svt_full_distortion_kernel_cbf_zero32_bits_cuda(int32_t *coeff, uint32_t coeff_stride, uint64_t distortion_result[2],
                                                     uint32_t area_width, uint32_t area_height) {
    uint64_t r[2];
    reduce2(coeff, coeff_stride, nullptr, 0, area_width, area_height, 4, M_SQ32, r);
    distortion_result[0] = distortion_result[1] = r[0];
}
uint64_t svt_spatial_full_distortion_kernel_cuda(uint8_t *input, uint32_t input_offset, uint32_t input_stride, uint8_t *recon,
                                                 int32_t recon_offset, uint32_t recon_stride, uint32_t area_width,
                                                 uint32_t area_height) {
    uint64_t r[2];
    reduce2(input + input_offset, input_stride, recon + recon_offset, recon_stride, area_width, area_height, 1, M_SSE8, r);
    return r[0];
}
uint64_t svt_full_distortion_kernel16_bits_cuda(uint8_t *input, uint32_t input_offset, uint32_t input_stride, uint8_t *recon,
                                                int32_t recon_offset, uint32_t recon_stride, uint32_t area_width,
                                                uint32_t area_height) {
    uint64_t r[2];
    reduce2((uint16_t *)input + input_offset, input_stride, (uint16_t *)recon + recon_offset, recon_stride, area_width, area_height, 2,
            M_SSE16, r);
    return r[0];
}
int svt_aom_satd_cuda(const int32_t *coeff, int length) {
    uint64_t r[2];
    reduce2(coeff, length, nullptr, 0, length, 1, 4, M_SATD, r);
    return (int)r[0];
}
int64_t svt_av1_block_error_cuda(const int32_t *coeff, const int32_t *dqcoeff, intptr_t block_size, int64_t *ssz) {
    uint64_t r[2];
    reduce2(coeff, block_size, dqcoeff, block_size, (int)block_size, 1, 4, M_BLKERR, r);
    *ssz = (int64_t)r[1];
    return (int64_t)r[0];
}
uint32_t svt_nxm_sad_kernel_sub_sampled_cuda(const uint8_t *src, uint32_t src_stride, const uint8_t *ref, uint32_t ref_stride,
                                             uint32_t height, uint32_t width) {
    uint64_t r[2]; // the C table entry is svt_nxm_sad_kernel_helper_c (aom_dsp_rtcd.c:372): a plain N x M SAD
    reduce2(src, src_stride, ref, ref_stride, width, height, 1, M_SAD8, r);
    return (uint32_t)r[0];
}
uint32_t sad_16b_kernel_cuda(uint16_t *src, uint32_t src_stride, uint16_t *ref, uint32_t ref_stride, uint32_t height, uint32_t width) {
    uint64_t r[2];
    reduce2(src, src_stride, ref, ref_stride, width, height, 2, M_SAD16, r);
    return (uint32_t)r[0];
}

#define SAD_MXN(W, H)                                                                                                   \
    uint32_t svt_aom_sad##W##x##H##_cuda(const uint8_t *src_ptr, int src_stride, const uint8_t *ref_ptr, int ref_stride) { \
        uint64_t r[2];                                                                                                  \
        reduce2(src_ptr, src_stride, ref_ptr, ref_stride, W, H, 1, M_SAD8, r);                                          \
        return (uint32_t)r[0];                                                                                          \
    }                                                                                                                   \
    void svt_aom_sad##W##x##H##x4d_cuda(const uint8_t *src_ptr, int src_stride, const uint8_t *const ref_ptr[],          \
                                        int ref_stride, uint32_t *sad_array) {                                          \
        for (int i = 0; i < 4; i++) sad_array[i] = svt_aom_sad##W##x##H##_cuda(src_ptr, src_stride, ref_ptr[i], ref_stride); \
    }
SAD_MXN(128, 128) SAD_MXN(128, 64) SAD_MXN(64, 128) SAD_MXN(64, 64) SAD_MXN(64, 32) SAD_MXN(64, 16) SAD_MXN(32, 64)
SAD_MXN(32, 32) SAD_MXN(32, 16) SAD_MXN(32, 8) SAD_MXN(16, 64) SAD_MXN(16, 32) SAD_MXN(16, 16) SAD_MXN(16, 8)
SAD_MXN(16, 4) SAD_MXN(8, 32) SAD_MXN(8, 16) SAD_MXN(8, 8) SAD_MXN(8, 4) SAD_MXN(4, 16) SAD_MXN(4, 8) SAD_MXN(4, 4)
}
