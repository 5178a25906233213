// subpel.cu — sub-pel motion refinement.
//
// Replaces, for a batch of independent (block, reference) searches, svt_av1_find_best_sub_pixel_tree
// (Source/Lib/Encoder/Codec/mcomp.c:350-418) as md_subpel_search sets it up (EbProductCodingLoop.c:2063-2155):
//   svt_upsampled_setup_center_error :323, svt_first_level_check :181, svt_second_level_check_v2 :261, svt_check_better :147,
//   svt_mv_err_cost :44, svt_aom_upsampled_pred_c (Encoder/C_DEFAULT/variance.c:212-269: two passes of
//   svt_aom_convolve8_horiz/vert, each rounded and clipped to 8 bits), svt_aom_varianceWxH_c (EbComputeVariance_C.c:14-61).
//
// One CTA per job; the candidates of a round that do not depend on each other are filtered as one batch (eval_batch).
// The search never leaves start_mv +- 14/8 sample (each of the three rounds moves at most twice its
// step per axis), so the (h + 10) x (w + 10) reference window and the source block are staged in shared memory ONCE and
// every one of the <= 25 candidate evaluations runs out of shared memory:
//   pass 1  (row, 4 columns) per thread: 4 aligned LDS.32, funnel-shift realignment, 2 dp4a per output with the AV1
//           kernels halved into int8 (all taps are even: clip8((2S + 64) >> 7) == clip8((S + 32) >> 6)); the 8-bit
//           result is stored TRANSPOSED so that
//   pass 2  (column, 4 rows) per thread reads its 11 vertical neighbours as 3 aligned words, 2 dp4a per output again, and
//           accumulates sum / sum of squares of (pred - src); warp shuffles + one shared-memory hop reduce them.
// A phase of 0 uses the identity kernel {0,0,0,128,...}, which reproduces the sample exactly, so the reference's four
// cases (copy / horizontal only / vertical only / both) are one code path.
// The decision logic (cost comparison order, diagonal choice, second-level rules) is executed redundantly by every thread
// on the broadcast (variance, sse) pair. Bound: instruction issue / shared memory (each sample is filtered ~25 times);
// HBM traffic is the window + block once per job.
#include <algorithm>

#include "common.cuh"

namespace svtb200 {
namespace {

// half taps (see interp.cu): regular 8-tap, regular 4-tap; bilinear is computed
__constant__ int8_t c_sub_taps[2][8][8]; // [8-tap | 4-tap][1/8 phase][tap]
const int8_t h_sub_taps[2][8][8] = {{{0, 0, 0, 64, 0, 0, 0, 0},
                                     {0, 1, -5, 61, 9, -2, 0, 0},
                                     {0, 1, -7, 55, 19, -5, 1, 0},
                                     {0, 1, -8, 47, 29, -6, 1, 0},
                                     {0, 1, -7, 38, 38, -7, 1, 0},
                                     {0, 1, -6, 29, 47, -8, 1, 0},
                                     {0, 1, -5, 19, 55, -7, 1, 0},
                                     {0, 0, -2, 9, 61, -5, 1, 0}},
                                    {{0, 0, 0, 64, 0, 0, 0, 0},
                                     {0, 0, -4, 61, 9, -2, 0, 0},
                                     {0, 0, -6, 55, 19, -4, 0, 0},
                                     {0, 0, -7, 47, 29, -5, 0, 0},
                                     {0, 0, -6, 38, 38, -6, 0, 0},
                                     {0, 0, -5, 29, 47, -7, 0, 0},
                                     {0, 0, -4, 19, 55, -6, 0, 0},
                                     {0, 0, -2, 9, 61, -4, 0, 0}}};

struct SubpelDev {
    SvtB200SubpelParams p;
    const uint8_t *src;
    int src_stride;
    const uint8_t *ref[8];
    int ref_stride[8];
    const SvtB200SubpelJob *jobs;
    SvtB200SubpelResult *results;
    int n_jobs, max_w, max_h, n_refs;
};

constexpr int SP_NT_MAX = 128; // CTA size is a template parameter: 32 / 64 / 128 threads by the largest block of the batch
constexpr int REACH = 2; // whole samples the search can move up/left of the start position (14/8 -> -2 .. +1)

__device__ __forceinline__ int dp4a_us(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint2 sub_taps(int type, int phase) { // av1_get_filter + the (phase << 1) row, halved
    if (type == 1) return make_uint2((uint32_t)(64 - 8 * phase) << 24, (uint32_t)(8 * phase)); // bilinear: taps 3, 4
    const int2 v = *reinterpret_cast<const int2 *>(c_sub_taps[type == 3 ? 0 : 1][phase]);
    return make_uint2((uint32_t)v.x, (uint32_t)v.y);
}
__device__ __forceinline__ int clip8(int v) { return min(max(v, 0), 255); }

struct Geo {
    int w, h, wp, tp; // window pitch (bytes), transposed-intermediate pitch (bytes)
    uint32_t inv_h7, inv_w; // ceil(2^32 / (h + 7)), ceil(2^32 / w): exact quotients by __umulhi for the item counts here
    uint8_t *win, *tmp, *src;
};

struct Mv2 {
    int row, col;
};

constexpr int MAX_BATCH = 4; // candidates evaluated between two barriers (left, right, up, down of a round)

// svt_upsampled_pref_error for up to N candidates at once, by the whole CTA: the candidates of a round are independent of
// each other (they depend on the round's centre only), so their first passes run back to back into N intermediate
// buffers, one barrier, then their second passes, ONE reduction of the 2 N sums — 3 barriers per batch instead of 3 per
// candidate. out[c] = {variance, sum of squares}; inactive candidates (outside the MV limits) are skipped.
template <int SP_NT, int N>
__device__ __forceinline__ void eval_batch(const Geo &g, const SvtB200SubpelParams &p, Mv2 start, const Mv2 (&mv)[N], const bool (&act)[N],
                                           uint2 (&out)[N], int *s_red) {
    const int ngx = g.w >> 2;
    const int tmp_bytes = g.w * g.tp;
    // pass 1: rows 0 .. h + 6 (3 above, 4 below the block), 4 columns per thread, transposed 8-bit output
#pragma unroll
    for (int c = 0; c < N; c++) {
        if (!act[c]) continue;
        const int drow = (mv[c].row >> 3) - (start.row >> 3), dcol = (mv[c].col >> 3) - (start.col >> 3); // -REACH .. REACH - 1
        const uint2 tx = sub_taps(p.subpel_search_type, mv[c].col & 7);
        const int c0 = dcol + REACH; // byte offset of tap 0 of output column 0 in a window row (0..3)
        const int r0 = drow + REACH; // window row of tap 0 of output row 0
        uint8_t *tmp = g.tmp + c * tmp_bytes;
        for (int it = threadIdx.x; it < (g.h + 7) * ngx; it += SP_NT) {
            const int xg = (int)__umulhi((uint32_t)it, g.inv_h7), r = it - xg * (g.h + 7); // r fastest: contiguous transposed stores
            const uint32_t *wr = reinterpret_cast<const uint32_t *>(g.win + (r0 + r) * g.wp) + xg;
            const uint32_t a0 = wr[0], a1 = wr[1], a2 = wr[2], a3 = wr[3];
            const uint32_t W0 = __funnelshift_r(a0, a1, 8 * c0), W1 = __funnelshift_r(a1, a2, 8 * c0), W2 = __funnelshift_r(a2, a3, 8 * c0);
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const uint32_t q0 = e ? __funnelshift_r(W0, W1, 8 * e) : W0, q1 = e ? __funnelshift_r(W1, W2, 8 * e) : W1;
                const int s = dp4a_us(q1, tx.y, dp4a_us(q0, tx.x, 0));
                tmp[(4 * xg + e) * g.tp + r] = (uint8_t)clip8((s + 32) >> 6);
            }
        }
    }
    __syncthreads();
    // pass 2 + the variance terms: column x, rows 4 yg .. 4 yg + 3
    int sum[N];
    unsigned sse[N];
#pragma unroll
    for (int c = 0; c < N; c++) {
        sum[c] = 0, sse[c] = 0;
        if (!act[c]) continue;
        const uint2 ty = sub_taps(p.subpel_search_type, mv[c].row & 7);
        const uint8_t *tmp = g.tmp + c * tmp_bytes;
        for (int it = threadIdx.x; it < g.w * (g.h >> 2); it += SP_NT) {
            const int yg = (int)__umulhi((uint32_t)it, g.inv_w), x = it - yg * g.w;
            const uint32_t *tc = reinterpret_cast<const uint32_t *>(tmp + x * g.tp) + yg;
            const uint32_t W0 = tc[0], W1 = tc[1], W2 = tc[2];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const uint32_t q0 = e ? __funnelshift_r(W0, W1, 8 * e) : W0, q1 = e ? __funnelshift_r(W1, W2, 8 * e) : W1;
                const int pr = clip8((dp4a_us(q1, ty.y, dp4a_us(q0, ty.x, 0)) + 32) >> 6);
                const int diff = pr - (int)g.src[(4 * yg + e) * g.w + x];
                sum[c] += diff;
                sse[c] += (unsigned)(diff * diff);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < N; c++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum[c] += __shfl_xor_sync(0xffffffffu, sum[c], o);
            sse[c] += __shfl_xor_sync(0xffffffffu, sse[c], o);
        }
        if ((threadIdx.x & 31) == 0) {
            s_red[((threadIdx.x >> 5) * MAX_BATCH + c) * 2] = sum[c];
            s_red[((threadIdx.x >> 5) * MAX_BATCH + c) * 2 + 1] = (int)sse[c];
        }
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < N; c++) {
        int sm = 0;
        unsigned sq = 0;
#pragma unroll
        for (int wq = 0; wq < SP_NT / 32; wq++) {
            sm += s_red[(wq * MAX_BATCH + c) * 2];
            sq += (unsigned)s_red[(wq * MAX_BATCH + c) * 2 + 1];
        }
        out[c] = make_uint2(sq - (unsigned)(((long long)sm * sm) / (g.w * g.h)), sq);
    }
    __syncthreads(); // s_red and the intermediates are free again
}

__device__ int mv_err_cost(const SvtB200SubpelParams &p, const SvtB200SubpelJob &j, Mv2 mv) {
    const int dr = (int16_t)(mv.row - j.ref_mv_row), dc = (int16_t)(mv.col - j.ref_mv_col);
    switch (p.mv_cost_type) {
    case 0: {
        const int joint = dr == 0 ? (dc == 0 ? 0 : 1) : (dc == 0 ? 2 : 3);
        const long long rate = (long long)(p.mvjcost[joint] + __ldg(p.mvcost[0] + dr) + __ldg(p.mvcost[1] + dc));
        return (int)((rate * p.error_per_bit + (1ll << 13)) >> 14);
    }
    case 1: return (2 * (abs(dr) + abs(dc))) >> 3;
    case 3: return (abs(dr) + abs(dc)) >> 3;
    default: return 0;
    }
}

struct Best {
    Mv2 mv;
    unsigned besterr, sse;
    int distortion;
};

__device__ __forceinline__ bool in_limits(const SvtB200SubpelJob &j, Mv2 mv) { // svt_av1_is_subpelmv_in_range
    return mv.col >= j.col_min && mv.col <= j.col_max && mv.row >= j.row_min && mv.row <= j.row_max;
}

// svt_check_better for N candidates whose errors do not depend on each other: evaluated together, then the reference's
// updates of (besterr, best_mv, distortion, sse) applied in its order. cost[c] = what svt_check_better returns.
template <int SP_NT, int N>
__device__ __forceinline__ void check_better_batch(const Geo &g, const SvtB200SubpelParams &p, const SvtB200SubpelJob &j, Mv2 start,
                                                   const Mv2 (&mv)[N], Best &b, int &is_better, unsigned (&cost)[N], int *s_red) {
    bool act[N];
    unsigned rate[N];
    uint2 ev[N];
#pragma unroll
    for (int c = 0; c < N; c++) {
        act[c] = in_limits(j, mv[c]);
        rate[c] = act[c] ? (unsigned)mv_err_cost(p, j, mv[c]) : 0u; // issued first: the table reads overlap the filtering
    }
    eval_batch<SP_NT, N>(g, p, start, mv, act, ev, s_red);
#pragma unroll
    for (int c = 0; c < N; c++) {
        if (!act[c]) {
            cost[c] = 0x7fffffffu; // INT_MAX
            continue;
        }
        const int thismse = (int)ev[c].x;
        cost[c] = rate[c] + (unsigned)thismse;
        if (cost[c] < b.besterr) {
            b.besterr = cost[c];
            b.mv = mv[c];
            b.distortion = thismse;
            b.sse = ev[c].y;
            is_better |= 1;
        }
    }
}

template <int SP_NT>
__global__ void __launch_bounds__(SP_NT) subpel_kernel(const __grid_constant__ SubpelDev d) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_red[2 * MAX_BATCH * SP_NT_MAX / 32];
    const SvtB200SubpelJob j = d.jobs[blockIdx.x];
    const SvtB200SubpelParams &p = d.p;
    // rejected (besterr = -1): blocks that do not fit the window of this launch, a start MV that is not full-pel (the staged
    // window and the REACH arithmetic assume md_subpel_search's full-pel start, EbProductCodingLoop.c:2094) and a reference
    // index outside the pictures that were passed
    if (j.bw > d.max_w || j.bh > d.max_h || j.bw < 4 || j.bh < 4 || ((j.bw | j.bh) & 3) || ((j.start_mv_row | j.start_mv_col) & 7) ||
        j.ref >= d.n_refs) {
        if (threadIdx.x == 0) d.results[blockIdx.x] = SvtB200SubpelResult{j.start_mv_row, j.start_mv_col, -1, -1, 0u};
        return;
    }
    Geo g;
    g.w = j.bw, g.h = j.bh;
    g.wp = ((g.w + 2 * REACH + 7 + 4 + 3) & ~3) | 4;  // >= w + 11 + the realignment over-read, pitch / 4 odd
    g.tp = ((g.h + 8 + 3) & ~3) | 4;                   // >= h + 8, pitch / 4 odd
    g.inv_h7 = (uint32_t)(((1ull << 32) + g.h + 6) / (uint32_t)(g.h + 7));
    g.inv_w = (uint32_t)(((1ull << 32) + g.w - 1) / (uint32_t)g.w);
    g.win = smem;
    g.tmp = g.win + (g.h + 2 * REACH + 7) * g.wp;
    g.src = g.tmp + MAX_BATCH * g.w * g.tp;
    const Mv2 start{j.start_mv_row, j.start_mv_col};
    // stage the window (rows / columns -REACH - 3 .. of the start position) and the source block
    const uint8_t *ref = d.ref[j.ref & 7] + (ptrdiff_t)(j.blk_y + (start.row >> 3) - REACH - 3) * d.ref_stride[j.ref & 7] + j.blk_x +
                         (start.col >> 3) - REACH - 3;
    const int ww = g.w + 2 * REACH + 7, wh = g.h + 2 * REACH + 7;
    for (int i = threadIdx.x; i < wh * g.wp; i += SP_NT) {
        const int r = i / g.wp, c = i - r * g.wp;
        g.win[i] = c < ww ? __ldg(ref + (ptrdiff_t)r * d.ref_stride[j.ref & 7] + c) : 0;
    }
    const uint8_t *sp = d.src + (ptrdiff_t)j.blk_y * d.src_stride + j.blk_x;
    for (int i = threadIdx.x; i < g.w * g.h; i += SP_NT) {
        const int r = i / g.w, c = i - r * g.w;
        g.src[i] = __ldg(sp + (ptrdiff_t)r * d.src_stride + c);
    }
    __syncthreads();

    // svt_av1_find_best_sub_pixel_tree
    const int round = min(3 - p.forced_stop, 3 - (p.allow_hp ? 0 : 1));
    Best b;
    b.mv = start;
    {
        const unsigned rate0 = (unsigned)mv_err_cost(p, j, start);
        const Mv2 m1[1] = {start};
        const bool a1[1] = {true};
        uint2 e1[1];
        eval_batch<SP_NT, 1>(g, p, start, m1, a1, e1, s_red);
        b.besterr = e1[0].x, b.sse = e1[0].y;
        b.distortion = (int)b.besterr;
        b.besterr += rate0;
    }
    int hstep = 4;
    for (int iter = 0; iter < round; iter++) {
        const Mv2 ctr = b.mv;
        int dummy = 0;
        // svt_first_level_check: left, right, up, down (one batch), then the diagonal they point to
        const Mv2 m4[4] = {{ctr.row, ctr.col - hstep}, {ctr.row, ctr.col + hstep}, {ctr.row - hstep, ctr.col}, {ctr.row + hstep, ctr.col}};
        unsigned c4[4];
        check_better_batch<SP_NT, 4>(g, p, j, start, m4, b, dummy, c4, s_red);
        Mv2 diag{c4[2] <= c4[3] ? -hstep : hstep, c4[0] <= c4[1] ? -hstep : hstep};
        const Mv2 md[1] = {{ctr.row + diag.row, ctr.col + diag.col}};
        unsigned cd[1];
        check_better_batch<SP_NT, 1>(g, p, j, start, md, b, dummy, cd, s_red);
        if (!(ctr.row == b.mv.row && ctr.col == b.mv.col) && p.iters_per_step > 1) { // svt_second_level_check_v2
            if (ctr.row == b.mv.row)
                diag.row = -diag.row;
            else if (ctr.col == b.mv.col)
                diag.col = -diag.col;
            const Mv2 m2[2] = {{b.mv.row + diag.row, b.mv.col}, {b.mv.row, b.mv.col + diag.col}};
            const Mv2 mb[1] = {{b.mv.row + diag.row, b.mv.col + diag.col}};
            int has_better = 0;
            unsigned c2[2];
            check_better_batch<SP_NT, 2>(g, p, j, start, m2, b, has_better, c2, s_red);
            if (has_better) check_better_batch<SP_NT, 1>(g, p, j, start, mb, b, has_better, cd, s_red);
        }
        hstep >>= 1;
    }
    if (threadIdx.x == 0) {
        SvtB200SubpelResult r;
        r.mv_row = (int16_t)b.mv.row, r.mv_col = (int16_t)b.mv.col;
        r.besterr = (int32_t)b.besterr, r.distortion = b.distortion, r.sse = b.sse;
        d.results[blockIdx.x] = r;
    }
}

size_t smem_bytes(int w, int h) {
    const int wp = ((w + 2 * REACH + 7 + 4 + 3) & ~3) | 4, tp = ((h + 8 + 3) & ~3) | 4;
    return (size_t)(h + 2 * REACH + 7) * wp + (size_t)MAX_BATCH * w * tp + (size_t)w * h + 16;
}

} // namespace
} // namespace svtb200

using namespace svtb200;

extern "C" int svt_b200_subpel_search(const SvtB200SubpelParams *p, const SvtB200Frame *src, const SvtB200Frame *refs,
                                      int32_t n_ref_frames, const SvtB200SubpelJob *jobs, int32_t n_jobs, SvtB200SubpelResult *results,
                                      void *stream) {
    if (!p || !src || !refs || n_ref_frames < 1 || n_ref_frames > 8 || n_jobs < 0 || (n_jobs && (!jobs || !results)) ||
        src->bit_depth != 8 || p->subpel_search_type < 1 || p->subpel_search_type > 3 || p->forced_stop < 0 || p->forced_stop > 3 ||
        p->mv_cost_type < 0 || p->mv_cost_type > 4 || (p->mv_cost_type == 0 && (!p->mvcost[0] || !p->mvcost[1])) ||
        (p->max_block_w && (p->max_block_w < 4 || p->max_block_w > 128)) || (p->max_block_h && (p->max_block_h < 4 || p->max_block_h > 128))) {
        set_error("svt_b200_subpel_search: bad argument (8-bit, 1..8 references, search type 1..3, cost tables for MV_COST_ENTROPY)");
        return SVT_B200_ERR_ARG;
    }
    if (n_jobs == 0) return SVT_B200_OK;
    static std::atomic<int> tables_done[64];
    int dev = 0;
    SVTB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && !tables_done[dev].load()) {
        SVTB_CUDA_TRY(cudaMemcpyToSymbol(c_sub_taps, h_sub_taps, sizeof(h_sub_taps)));
        SVTB_CUDA_TRY(cudaFuncSetAttribute(subpel_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(128, 128)));
        tables_done[dev].store(1);
    }
    SubpelDev d;
    memset(&d, 0, sizeof(d));
    d.p = *p;
    d.src = (const uint8_t *)src->y;
    d.src_stride = src->stride_y;
    d.n_refs = n_ref_frames;
    for (int i = 0; i < n_ref_frames; i++) {
        if (refs[i].bit_depth != 8) {
            set_error("svt_b200_subpel_search: reference %d is not 8-bit", i);
            return SVT_B200_ERR_ARG;
        }
        d.ref[i] = (const uint8_t *)refs[i].y;
        d.ref_stride[i] = refs[i].stride_y;
    }
    d.jobs = jobs, d.results = results, d.n_jobs = n_jobs;
    const int mw = p->max_block_w ? p->max_block_w : 128, mh = p->max_block_h ? p->max_block_h : 128;
    const size_t smem = smem_bytes(mw, mh);
    d.max_w = mw, d.max_h = mh;
    cudaStream_t st = (cudaStream_t)stream;
    if (mw * mh <= 64) // a pass is (h + 7) x w / 4 resp. w x h / 4 work items: one warp covers an 8x8 block
        SVTB_LAUNCH(subpel_kernel<32>, n_jobs, 32, smem, st, d);
    else if (mw * mh <= 1024)
        SVTB_LAUNCH(subpel_kernel<64>, n_jobs, 64, smem, st, d);
    else
        SVTB_LAUNCH(subpel_kernel<128>, n_jobs, 128, smem, st, d);
    SVTB_CUDA_TRY(cudaGetLastError());
    return SVT_B200_OK;
}
