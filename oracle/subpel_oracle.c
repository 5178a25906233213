/*
 * subpel_oracle.c — CPU restatement of the sub-pel motion refinement.
 *
 * TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this file; the
 * product (libsvtav1_b200.so) never links or calls it.
 *
 * Follows, in svt-av1 v0.8.6:
 *   Source/Lib/Encoder/Codec/mcomp.c
 *     :44-64    svt_mv_err_cost            :102-143  svt_upsampled_pref_error
 *     :147-170  svt_check_better           :172-179  svt_get_best_diag_step
 *     :181-259  svt_first_level_check      :261-321  svt_second_level_check_v2
 *     :323-335  svt_upsampled_setup_center_error     :350-418 svt_av1_find_best_sub_pixel_tree
 *   Source/Lib/Encoder/Codec/mcomp.h:140-152  svt_av1_is_subpelmv_in_range, svt_mv_cost
 *   Source/Lib/Encoder/C_DEFAULT/variance.c:200-269  av1_get_filter, svt_aom_upsampled_pred_c
 *   Source/Lib/Common/Codec/convolve.c:249-308        svt_aom_convolve8_horiz_c / _vert_c (step 16)
 *   Source/Lib/Encoder/C_DEFAULT/EbComputeVariance_C.c:14-61  variance_c, svt_aom_varianceWxH_c
 * tests/test_oracle_subpel.py pins orc_subpel_search against svt_av1_find_best_sub_pixel_tree itself (oracle/_ref).
 */
#include <limits.h>
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

typedef struct {
    int row, col;
} Mv;
typedef struct {
    const SvtB200SubpelParams *p;
    const SvtB200SubpelJob *j;
    const uint8_t *src, *ref; /* sample (0,0) of the block in the source / reference luma plane */
    int src_stride, ref_stride;
    const int32_t *mvcost[2]; /* host copies of the centred tables */
} Ctx;

static inline int clip8(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }

/* svt_aom_upsampled_pred_c + the variance of (pred - src) */
static unsigned eval_error(const Ctx *c, Mv mv, unsigned *sse_out) {
    const int w = c->j->bw, h = c->j->bh, rs = c->ref_stride;
    const uint8_t *ref = c->ref + (mv.row >> 3) * rs + (mv.col >> 3); /* svt_get_buf_from_mv */
    const int sx = mv.col & 7, sy = mv.row & 7;
    /* av1_get_filter: 2 taps -> bilinear, 4 taps -> the regular 4-tap kernel, 8 taps -> regular; row = (1/8 phase) << 1 */
    const int t = c->p->subpel_search_type;
    int16_t kx[8], ky[8];
    orc_interp_kernel(t == 1 ? 3 : 0, t == 3 ? 8 : 4, sx << 1, kx);
    orc_interp_kernel(t == 1 ? 3 : 0, t == 3 ? 8 : 4, sy << 1, ky);
    uint8_t *temp = NULL;
    if (sx && sy) { /* first pass: h + 7 rows from 3 above, rounded and clipped to 8 bits */
        temp = (uint8_t *)malloc((size_t)(h + 7) * w);
        for (int y = 0; y < h + 7; y++)
            for (int x = 0; x < w; x++) {
                int sum = 0;
                for (int k = 0; k < 8; k++) sum += kx[k] * ref[(y - 3) * rs + x - 3 + k];
                temp[y * w + x] = (uint8_t)clip8((sum + 64) >> 7);
            }
    }
    int sum = 0;
    unsigned sse = 0;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int pr;
            if (!sx && !sy)
                pr = ref[y * rs + x];
            else if (!sy) {
                int s = 0;
                for (int k = 0; k < 8; k++) s += kx[k] * ref[y * rs + x - 3 + k];
                pr = clip8((s + 64) >> 7);
            } else if (!sx) {
                int s = 0;
                for (int k = 0; k < 8; k++) s += ky[k] * ref[(y - 3 + k) * rs + x];
                pr = clip8((s + 64) >> 7);
            } else {
                int s = 0;
                for (int k = 0; k < 8; k++) s += ky[k] * temp[(y + k) * w + x];
                pr = clip8((s + 64) >> 7);
            }
            const int diff = pr - c->src[y * c->src_stride + x];
            sum += diff;
            sse += (unsigned)(diff * diff);
        }
    free(temp);
    *sse_out = sse;
    return sse - (unsigned)(((int64_t)sum * sum) / (w * h));
}

/* svt_aom_upsampled_pred_c (variance.c:212-269) on its own: comp_pred is width x height, packed */
ORC_API void orc_upsampled_pred(uint8_t *comp_pred, int width, int height, int subpel_x_q3, int subpel_y_q3, const uint8_t *ref,
                                int ref_stride, int subpel_search) {
    int16_t kx[8], ky[8];
    orc_interp_kernel(subpel_search == 1 ? 3 : 0, subpel_search == 3 ? 8 : 4, subpel_x_q3 << 1, kx);
    orc_interp_kernel(subpel_search == 1 ? 3 : 0, subpel_search == 3 ? 8 : 4, subpel_y_q3 << 1, ky);
    uint8_t *temp = (uint8_t *)malloc((size_t)(height + 7) * width);
    if (subpel_x_q3 && subpel_y_q3)
        for (int y = 0; y < height + 7; y++)
            for (int x = 0; x < width; x++) {
                int s = 0;
                for (int k = 0; k < 8; k++) s += kx[k] * ref[(y - 3) * ref_stride + x - 3 + k];
                temp[y * width + x] = (uint8_t)clip8((s + 64) >> 7);
            }
    for (int y = 0; y < height; y++)
        for (int x = 0; x < width; x++) {
            int s = 0, pr;
            if (!subpel_x_q3 && !subpel_y_q3)
                pr = ref[y * ref_stride + x];
            else {
                for (int k = 0; k < 8; k++)
                    s += !subpel_y_q3 ? kx[k] * ref[y * ref_stride + x - 3 + k]
                                      : !subpel_x_q3 ? ky[k] * ref[(y - 3 + k) * ref_stride + x] : ky[k] * temp[(y + k) * width + x];
                pr = clip8((s + 64) >> 7);
            }
            comp_pred[y * width + x] = (uint8_t)pr;
        }
    free(temp);
}

static int mv_err_cost(const Ctx *c, Mv mv) {
    const int16_t dr = (int16_t)(mv.row - c->j->ref_mv_row), dc = (int16_t)(mv.col - c->j->ref_mv_col);
    const int ar = abs(dr), ac = abs(dc);
    switch (c->p->mv_cost_type) {
    case 0: { /* MV_COST_ENTROPY: joint + component rates, scaled by error_per_bit */
        const int joint = dr == 0 ? (dc == 0 ? 0 : 1) : (dc == 0 ? 2 : 3);
        const int64_t rate = c->p->mvjcost[joint] + c->mvcost[0][dr] + c->mvcost[1][dc];
        return (int)((rate * c->p->error_per_bit + ((int64_t)1 << 13)) >> 14);
    }
    case 1: return (2 * (ar + ac)) >> 3;
    case 2: return 0;
    case 3: return (1 * (ar + ac)) >> 3;
    default: return 0;
    }
}

typedef struct {
    Mv best;
    unsigned besterr, sse;
    int distortion;
} Best;

static unsigned check_better(const Ctx *c, Mv mv, Best *b, int *is_better) {
    if (mv.col < c->j->col_min || mv.col > c->j->col_max || mv.row < c->j->row_min || mv.row > c->j->row_max) return INT_MAX;
    unsigned sse;
    const int thismse = (int)eval_error(c, mv, &sse);
    unsigned cost = (unsigned)mv_err_cost(c, mv);
    cost += (unsigned)thismse;
    if (cost < b->besterr) {
        b->besterr = cost;
        b->best = mv;
        b->distortion = thismse;
        b->sse = sse;
        *is_better |= 1;
    }
    return cost;
}

static void search_one(const Ctx *c, SvtB200SubpelResult *out) {
    const SvtB200SubpelParams *p = c->p;
    const int r_a = 3 - p->forced_stop, r_b = 3 - !p->allow_hp;
    const int round = r_a < r_b ? r_a : r_b;
    Best b;
    b.best.row = c->j->start_mv_row;
    b.best.col = c->j->start_mv_col;
    b.besterr = eval_error(c, b.best, &b.sse);
    b.distortion = (int)b.besterr;
    b.besterr += (unsigned)mv_err_cost(c, b.best);
    int hstep = 4;
    for (int iter = 0; iter < round; iter++) {
        const Mv ctr = b.best;
        int dummy = 0;
        const Mv l = {ctr.row, ctr.col - hstep}, r = {ctr.row, ctr.col + hstep}, u = {ctr.row - hstep, ctr.col},
                 d = {ctr.row + hstep, ctr.col};
        const unsigned left = check_better(c, l, &b, &dummy), right = check_better(c, r, &b, &dummy);
        const unsigned up = check_better(c, u, &b, &dummy), down = check_better(c, d, &b, &dummy);
        Mv diag = {up <= down ? -hstep : hstep, left <= right ? -hstep : hstep};
        const Mv dm = {ctr.row + diag.row, ctr.col + diag.col};
        check_better(c, dm, &b, &dummy);
        if (!(ctr.row == b.best.row && ctr.col == b.best.col) && p->iters_per_step > 1) { /* svt_second_level_check_v2 */
            if (ctr.row == b.best.row)
                diag.row *= -1;
            else if (ctr.col == b.best.col)
                diag.col *= -1;
            const Mv rb = {b.best.row + diag.row, b.best.col}, cb = {b.best.row, b.best.col + diag.col},
                     db = {b.best.row + diag.row, b.best.col + diag.col};
            int has_better = 0;
            check_better(c, rb, &b, &has_better);
            check_better(c, cb, &b, &has_better);
            if (has_better) check_better(c, db, &b, &has_better);
        }
        hstep >>= 1;
    }
    out->mv_row = (int16_t)b.best.row;
    out->mv_col = (int16_t)b.best.col;
    out->besterr = (int32_t)b.besterr;
    out->distortion = b.distortion;
    out->sse = b.sse;
}

/* mvcost0 / mvcost1: HOST copies of the centred tables (p->mvcost holds the device pointers of the product call) */
ORC_API void orc_subpel_search(const SvtB200SubpelParams *p, const int32_t *mvcost0, const int32_t *mvcost1, const SvtB200Frame *src,
                               const SvtB200Frame *refs, int n_ref_frames, const SvtB200SubpelJob *jobs, int n_jobs,
                               SvtB200SubpelResult *results) {
    (void)n_ref_frames;
#pragma omp parallel for schedule(dynamic, 8)
    for (int i = 0; i < n_jobs; i++) {
        Ctx c;
        c.p = p;
        c.j = &jobs[i];
        c.src_stride = src->stride_y;
        c.src = (const uint8_t *)src->y + (ptrdiff_t)jobs[i].blk_y * src->stride_y + jobs[i].blk_x;
        const SvtB200Frame *rf = &refs[jobs[i].ref];
        c.ref_stride = rf->stride_y;
        c.ref = (const uint8_t *)rf->y + (ptrdiff_t)jobs[i].blk_y * rf->stride_y + jobs[i].blk_x;
        c.mvcost[0] = mvcost0;
        c.mvcost[1] = mvcost1;
        search_one(&c, &results[i]);
    }
}
