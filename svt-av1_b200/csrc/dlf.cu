// dlf.cu — AV1 deblocking loop filter on sm_100a.
//
// Replaces (reference files under Source/Lib):
//   svt_aom_lpf_{horizontal,vertical}_{4,6,8,14}_c + svt_aom_highbd_lpf_*   Common/Codec/EbDeblockingCommon.c:251-924
//   set_lpf_parameters, svt_av1_filter_block_plane_vert/horz, loop_filter_sb,
//   svt_av1_loop_filter_frame                                              Encoder/Codec/EbDeblockingFilter.c:168-753
//   picture_sse_calculations (distortion of svt_av1_pick_filter_level)      Encoder/Codec/EbDeblockingFilter.c:830-964
//
// Design: the reference filters one 64x64 SB at a time on a single thread per picture (vertical edges of SB n,
// then horizontal edges of SB n-1).  AV1 guarantees that edges of one direction never read what another edge of
// the same direction writes and that no horizontal-edge filter touches a sample a later vertical-edge filter reads,
// so the whole picture is two launches: every vertical edge in parallel, then every horizontal edge — one thread
// per sample line of an edge, edge parameters (length, level) derived on the fly from a 16-byte-per-4x4 summary of
// the mode-info grid.  HBM traffic: the picture is read and written once per pass (2 B/sample + 1 B/edge, §8d).
#include <algorithm>

#include "common.cuh"

using namespace svtb200;

namespace {

__device__ __forceinline__ int sclamp(int v, int bd) {
    const int lo = -(128 << (bd - 8)), hi = (128 << (bd - 8)) - 1;
    return v < lo ? lo : (v > hi ? hi : v);
}

// One sample line of an edge. px: p6..p0 at [0..6], q0..q6 at [7..13]. Returns the number of taps per side that may
// have changed (0 = untouched).  Same decisions as filter4/6/8/14 and their masks (EbDeblockingCommon.c:148-924).
__device__ __forceinline__ int lpf_sample(int *px, int len, int blimit, int limit, int thresh, int bd) {
    const int sh = bd - 8;
#define P(i) px[6 - (i)]
#define Q(i) px[7 + (i)]
    const int lim = limit << sh, blim = blimit << sh, thr = thresh << sh, one = 1 << sh;
    bool mask = abs(P(1) - P(0)) <= lim && abs(Q(1) - Q(0)) <= lim && abs(P(0) - Q(0)) * 2 + abs(P(1) - Q(1)) / 2 <= blim;
    if (len >= 6) mask = mask && abs(P(2) - P(1)) <= lim && abs(Q(2) - Q(1)) <= lim;
    if (len >= 8) mask = mask && abs(P(3) - P(2)) <= lim && abs(Q(3) - Q(2)) <= lim;
    bool flat = false, flat2 = false;
    if (len >= 6) {
        flat = abs(P(1) - P(0)) <= one && abs(Q(1) - Q(0)) <= one && abs(P(2) - P(0)) <= one && abs(Q(2) - Q(0)) <= one;
        if (len >= 8) flat = flat && abs(P(3) - P(0)) <= one && abs(Q(3) - Q(0)) <= one;
    }
    if (len == 14)
        flat2 = abs(P(4) - P(0)) <= one && abs(Q(4) - Q(0)) <= one && abs(P(5) - P(0)) <= one && abs(Q(5) - Q(0)) <= one &&
            abs(P(6) - P(0)) <= one && abs(Q(6) - Q(0)) <= one;
    const int p0 = P(0), p1 = P(1), p2 = P(2), p3 = P(3), p4 = P(4), p5 = P(5), p6 = P(6);
    const int q0 = Q(0), q1 = Q(1), q2 = Q(2), q3 = Q(3), q4 = Q(4), q5 = Q(5), q6 = Q(6);
#define R(v, n) (((v) + (1 << ((n)-1))) >> (n))
    if (len == 14 && flat2 && flat && mask) {
        P(5) = R(p6 * 7 + p5 * 2 + p4 * 2 + p3 + p2 + p1 + p0 + q0, 4);
        P(4) = R(p6 * 5 + p5 * 2 + p4 * 2 + p3 * 2 + p2 + p1 + p0 + q0 + q1, 4);
        P(3) = R(p6 * 4 + p5 + p4 * 2 + p3 * 2 + p2 * 2 + p1 + p0 + q0 + q1 + q2, 4);
        P(2) = R(p6 * 3 + p5 + p4 + p3 * 2 + p2 * 2 + p1 * 2 + p0 + q0 + q1 + q2 + q3, 4);
        P(1) = R(p6 * 2 + p5 + p4 + p3 + p2 * 2 + p1 * 2 + p0 * 2 + q0 + q1 + q2 + q3 + q4, 4);
        P(0) = R(p6 + p5 + p4 + p3 + p2 + p1 * 2 + p0 * 2 + q0 * 2 + q1 + q2 + q3 + q4 + q5, 4);
        Q(0) = R(p5 + p4 + p3 + p2 + p1 + p0 * 2 + q0 * 2 + q1 * 2 + q2 + q3 + q4 + q5 + q6, 4);
        Q(1) = R(p4 + p3 + p2 + p1 + p0 + q0 * 2 + q1 * 2 + q2 * 2 + q3 + q4 + q5 + q6 * 2, 4);
        Q(2) = R(p3 + p2 + p1 + p0 + q0 + q1 * 2 + q2 * 2 + q3 * 2 + q4 + q5 + q6 * 3, 4);
        Q(3) = R(p2 + p1 + p0 + q0 + q1 + q2 * 2 + q3 * 2 + q4 * 2 + q5 + q6 * 4, 4);
        Q(4) = R(p1 + p0 + q0 + q1 + q2 + q3 * 2 + q4 * 2 + q5 * 2 + q6 * 5, 4);
        Q(5) = R(p0 + q0 + q1 + q2 + q3 + q4 * 2 + q5 * 2 + q6 * 7, 4);
        return 6;
    }
    if (len >= 8 && flat && mask) {
        P(2) = R(p3 + p3 + p3 + 2 * p2 + p1 + p0 + q0, 3);
        P(1) = R(p3 + p3 + p2 + 2 * p1 + p0 + q0 + q1, 3);
        P(0) = R(p3 + p2 + p1 + 2 * p0 + q0 + q1 + q2, 3);
        Q(0) = R(p2 + p1 + p0 + 2 * q0 + q1 + q2 + q3, 3);
        Q(1) = R(p1 + p0 + q0 + 2 * q1 + q2 + q3 + q3, 3);
        Q(2) = R(p0 + q0 + q1 + 2 * q2 + q3 + q3 + q3, 3);
        return 3;
    }
    if (len == 6 && flat && mask) {
        P(1) = R(p2 * 3 + p1 * 2 + p0 * 2 + q0, 3);
        P(0) = R(p2 + p1 * 2 + p0 * 2 + q0 * 2 + q1, 3);
        Q(0) = R(p1 + p0 * 2 + q0 * 2 + q1 * 2 + q2, 3);
        Q(1) = R(p0 + q0 * 2 + q1 * 2 + q2 * 3, 3);
        return 2;
    }
    if (!mask) return 0; // filter4 with mask = 0 is the identity
    const int off = 0x80 << sh;
    const int ps1 = p1 - off, ps0 = p0 - off, qs0 = q0 - off, qs1 = q1 - off;
    const bool hev = abs(p1 - p0) > thr || abs(q1 - q0) > thr;
    int f = hev ? sclamp(ps1 - qs1, bd) : 0;
    f = sclamp(f + 3 * (qs0 - ps0), bd);
    const int f1 = sclamp(f + 4, bd) >> 3, f2 = sclamp(f + 3, bd) >> 3;
    Q(0) = sclamp(qs0 - f1, bd) + off;
    P(0) = sclamp(ps0 + f2, bd) + off;
    const int f3 = hev ? 0 : ((f1 + 1) >> 1);
    Q(1) = sclamp(qs1 - f3, bd) + off;
    P(1) = sclamp(ps1 + f3, bd) + off;
    return 2;
#undef P
#undef Q
#undef R
}

__device__ __forceinline__ int taps_of(int len) { return len == 4 ? 2 : len == 6 ? 3 : len == 8 ? 4 : 7; }

struct DlfDev {
    SvtB200DlfParams p;
    void *plane[3];
    int stride[3];
    int bd;
    const SvtB200DlfMi *mi;
};

// set_lpf_parameters on the flattened summary
__device__ __forceinline__ int edge_params(const DlfDev &d, int plane, int vert, int x, int y, int &level) {
    const SvtB200DlfParams &p = d.p;
    const int ss = plane ? 1 : 0;
    const int mi_row = ss | ((y << ss) >> 2), mi_col = ss | ((x << ss) >> 2);
    const SvtB200DlfMi *cur = d.mi + (size_t)mi_row * p.mi_stride + mi_col;
    const int ts = vert ? cur->tx_w[ss] : cur->tx_h[ss];
    const int coord = vert ? x : y;
    if ((coord & (ts - 1)) || !coord) return 0;
    const SvtB200DlfMi *prev = vert ? cur - (1 << ss) : cur - (size_t)(1 << ss) * p.mi_stride;
    const int pv_ts = vert ? prev->tx_w[ss] : prev->tx_h[ss];
    const int cl = plane == 0 ? cur->lvl_y[vert ? 0 : 1] : plane == 1 ? cur->lvl_u : cur->lvl_v;
    const int pl = plane == 0 ? prev->lvl_y[vert ? 0 : 1] : plane == 1 ? prev->lvl_u : prev->lvl_v;
    const bool pu_edge = !(coord & ((vert ? cur->blk_w[ss] : cur->blk_h[ss]) - 1));
    if (!((cl || pl) && (!prev->skip_inter || !cur->skip_inter || pu_edge))) return 0;
    const int mn = min(ts, pv_ts);
    level = cl ? cl : pl;
    if (mn <= 4) return 4;
    if (mn == 8) return plane ? 6 : 8;
    return plane ? 6 : 14;
}
__device__ __forceinline__ void thresholds(int level, int sharpness, int &blimit, int &limit, int &thresh) {
    int lim = level >> ((sharpness > 0) + (sharpness > 4));
    if (sharpness > 0 && lim > 9 - sharpness) lim = 9 - sharpness;
    if (lim < 1) lim = 1;
    limit = lim;
    blimit = 2 * (level + 2) + lim;
    thresh = level >> 4;
}

// One thread per sample line of a 4x4 unit's leading edge.  vert = 1: thread = (row y, unit column ux);
// vert = 0: thread = (unit row uy, column x) so that a warp walks along a row of the picture (coalesced).
template <typename T>
__global__ void __launch_bounds__(256) dlf_pass_kernel(const __grid_constant__ DlfDev d, int planes, int vert) {
    const int plane = (planes >> (2 * blockIdx.z)) & 3; // the planes of one direction share a launch (grid.z)
    const int ss = plane ? 1 : 0;
    const int pw = (d.p.mi_cols * 4) >> ss, ph = (d.p.mi_rows * 4) >> ss;
    const int nx = vert ? pw / 4 : pw, ny = vert ? ph : ph / 4;
    const int gx = blockIdx.x * blockDim.x + threadIdx.x, gy = blockIdx.y;
    if (gx >= nx || gy >= ny) return;
    const int x = vert ? gx * 4 : gx, y = vert ? gy : gy * 4;
    int level = 0;
    const int len = edge_params(d, plane, vert, vert ? x : (x & ~3), vert ? (y & ~3) : y, level);
    if (!len) return;
    int bl, li, th;
    thresholds(level, d.p.sharpness, bl, li, th);
    T *s = reinterpret_cast<T *>(d.plane[plane]) + (size_t)y * d.stride[plane] + x;
    const ptrdiff_t across = vert ? 1 : d.stride[plane];
    const int n = taps_of(len);
    int px[14];
#pragma unroll
    for (int t = 0; t < 7; t++) {
        px[6 - t] = t < n ? (int)s[-(ptrdiff_t)(t + 1) * across] : 0;
        px[7 + t] = t < n ? (int)s[(ptrdiff_t)t * across] : 0;
    }
    const int changed = lpf_sample(px, len, bl, li, th, d.bd);
#pragma unroll
    for (int t = 0; t < 6; t++)
        if (t < changed) {
            s[-(ptrdiff_t)(t + 1) * across] = (T)px[6 - t];
            s[(ptrdiff_t)t * across] = (T)px[7 + t];
        }
}

// drop-in: one 4-line edge segment staged as a 16 x 4 window (p7..q7 across, 4 lines)
__global__ void lpf_edge_kernel(uint16_t *win, int len, int blimit, int limit, int thresh, int bd) {
    const int i = threadIdx.x;
    if (i >= 4) return;
    int px[14];
    for (int t = 0; t < 7; t++) {
        px[6 - t] = win[i * 16 + 7 - t];
        px[7 + t] = win[i * 16 + 8 + t];
    }
    lpf_sample(px, len, blimit, limit, thresh, bd);
    for (int t = 0; t < 7; t++) {
        win[i * 16 + 7 - t] = (uint16_t)px[6 - t];
        win[i * 16 + 8 + t] = (uint16_t)px[7 + t];
    }
}

template <typename T>
__global__ void __launch_bounds__(256) sse_kernel(const T *a, int sa, const T *b, int sb, int w, int h, unsigned long long *out) {
    unsigned long long acc = 0;
    for (int y = blockIdx.x; y < h; y += gridDim.x)
        for (int x = threadIdx.x; x < w; x += blockDim.x) {
            const int dlt = (int)a[(size_t)y * sa + x] - (int)b[(size_t)y * sb + x];
            acc += (unsigned long long)(dlt * dlt);
        }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

void lpf_dropin(void *s, int hbd, int pitch, int vert, int len, const uint8_t *blimit, const uint8_t *limit,
                const uint8_t *thresh, int bd) {
    ThreadCtx &c = tls();
    c.reserve(256);
    uint16_t *w = (uint16_t *)c.h;
    const int n = len == 4 ? 2 : len == 6 ? 3 : len == 8 ? 4 : 7;
    const ptrdiff_t across = vert ? 1 : pitch, along = vert ? pitch : 1;
    memset(w, 0, 128);
    for (int i = 0; i < 4; i++)
        for (int t = 0; t < n; t++) {
            const ptrdiff_t op = i * along - (t + 1) * across, oq = i * along + t * across;
            w[i * 16 + 7 - t] = hbd ? ((uint16_t *)s)[op] : ((uint8_t *)s)[op];
            w[i * 16 + 8 + t] = hbd ? ((uint16_t *)s)[oq] : ((uint8_t *)s)[oq];
        }
    SVTB_CUDA_FATAL(cudaMemcpyAsync(c.d, c.h, 128, cudaMemcpyHostToDevice, c.stream));
    SVTB_LAUNCH(lpf_edge_kernel, 1, 32, 0, c.stream, (uint16_t *)c.d, len, (int)*blimit, (int)*limit, (int)*thresh, bd);
    SVTB_CUDA_FATAL(cudaMemcpyAsync(c.h, c.d, 128, cudaMemcpyDeviceToHost, c.stream));
    SVTB_CUDA_FATAL(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < 4; i++)
        for (int t = 0; t < n; t++) {
            const ptrdiff_t op = i * along - (t + 1) * across, oq = i * along + t * across;
            if (hbd) {
                ((uint16_t *)s)[op] = w[i * 16 + 7 - t];
                ((uint16_t *)s)[oq] = w[i * 16 + 8 + t];
            } else {
                ((uint8_t *)s)[op] = (uint8_t)w[i * 16 + 7 - t];
                ((uint8_t *)s)[oq] = (uint8_t)w[i * 16 + 8 + t];
            }
        }
}

} // namespace

extern "C" {

#define LPF_DROPIN(DIR, VERT, N)                                                                                            \
    void svt_aom_lpf_##DIR##_##N##_cuda(uint8_t *s, int32_t pitch, const uint8_t *blimit, const uint8_t *limit,             \
                                        const uint8_t *thresh) {                                                            \
        lpf_dropin(s, 0, pitch, VERT, N, blimit, limit, thresh, 8);                                                         \
    }                                                                                                                       \
    void svt_aom_highbd_lpf_##DIR##_##N##_cuda(uint16_t *s, int32_t pitch, const uint8_t *blimit, const uint8_t *limit,     \
                                               const uint8_t *thresh, int32_t bd) {                                         \
        lpf_dropin(s, 1, pitch, VERT, N, blimit, limit, thresh, bd);                                                        \
    }
LPF_DROPIN(horizontal, 0, 4)
LPF_DROPIN(horizontal, 0, 6)
LPF_DROPIN(horizontal, 0, 8)
LPF_DROPIN(horizontal, 0, 14)
LPF_DROPIN(vertical, 1, 4)
LPF_DROPIN(vertical, 1, 6)
LPF_DROPIN(vertical, 1, 8)
LPF_DROPIN(vertical, 1, 14)

int svt_b200_dlf_frame(const SvtB200DlfParams *p, const SvtB200Frame *frame, const SvtB200DlfMi *mi, void *stream) {
    if (!p || !frame || !mi || !frame->y || !frame->cb || !frame->cr || p->mi_rows <= 0 || p->mi_cols <= 0 ||
        p->mi_stride < p->mi_cols) {
        set_error("svt_b200_dlf_frame: bad argument");
        return SVT_B200_ERR_ARG;
    }
    DlfDev d;
    d.p = *p;
    d.plane[0] = frame->y;
    d.plane[1] = frame->cb;
    d.plane[2] = frame->cr;
    d.stride[0] = frame->stride_y;
    d.stride[1] = d.stride[2] = frame->stride_c;
    d.bd = frame->bit_depth;
    d.mi = mi;
    cudaStream_t st = (cudaStream_t)stream;
    const bool hbd = frame->bit_depth > 8;
    // all vertical edges of every plane first, then all horizontal edges (see the header comment)
    int planes = 0, n_planes = 0, max_w = 0, max_h = 0;
    for (int plane = p->plane_start; plane < p->plane_end; plane++) {
        if (plane == 0 && !p->filter_level[0] && !p->filter_level[1]) break; // loop_filter_sb :629-636
        if (plane == 1 && !p->filter_level_u) continue;
        if (plane == 2 && !p->filter_level_v) continue;
        const int ss = plane ? 1 : 0;
        planes |= plane << (2 * n_planes++);
        max_w = std::max(max_w, (p->mi_cols * 4) >> ss);
        max_h = std::max(max_h, (p->mi_rows * 4) >> ss);
    }
    for (int vert = 1; vert >= 0 && n_planes; vert--) {
        const int nx = vert ? max_w / 4 : max_w, ny = vert ? max_h : max_h / 4;
        dim3 grid((nx + 255) / 256, ny, n_planes);
        if (hbd)
            SVTB_LAUNCH(dlf_pass_kernel<uint16_t>, grid, 256, 0, st, d, planes, vert);
        else
            SVTB_LAUNCH(dlf_pass_kernel<uint8_t>, grid, 256, 0, st, d, planes, vert);
    }
    SVTB_CUDA_TRY(cudaGetLastError());
    return SVT_B200_OK;
}

int svt_b200_frame_sse(const SvtB200Frame *a, const SvtB200Frame *b, uint64_t *sse, void *stream) {
    if (!a || !b || !sse || a->bit_depth != b->bit_depth || a->width != b->width || a->height != b->height) {
        set_error("svt_b200_frame_sse: bad argument");
        return SVT_B200_ERR_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    SVTB_CUDA_TRY(cudaMemsetAsync(sse, 0, 3 * sizeof(uint64_t), st));
    for (int pl = 0; pl < 3; pl++) {
        const int w = pl ? (a->width + 1) >> 1 : a->width, h = pl ? (a->height + 1) >> 1 : a->height;
        const void *pa = pl == 0 ? a->y : pl == 1 ? a->cb : a->cr, *pb = pl == 0 ? b->y : pl == 1 ? b->cb : b->cr;
        const int sa = pl ? a->stride_c : a->stride_y, sb = pl ? b->stride_c : b->stride_y;
        const int grid = h < 592 ? h : 592;
        if (a->bit_depth > 8)
            SVTB_LAUNCH(sse_kernel<uint16_t>, grid, 256, 0, st, (const uint16_t *)pa, sa, (const uint16_t *)pb, sb, w, h,
                        (unsigned long long *)(sse + pl));
        else
            SVTB_LAUNCH(sse_kernel<uint8_t>, grid, 256, 0, st, (const uint8_t *)pa, sa, (const uint8_t *)pb, sb, w, h,
                        (unsigned long long *)(sse + pl));
    }
    SVTB_CUDA_TRY(cudaGetLastError());
    return SVT_B200_OK;
}
}
