"""Sub-pel refinement test cases: a source picture, references that are sub-sample shifted / noisy versions of it, a batch
of (block, reference) searches with the limits md_subpel_search derives, and MV rate tables of a realistic shape."""
import ctypes as C

import numpy as np

import common as cm
import interp_cases as ic
import svtb200 as sb

MV_MAX = 16383
BLOCKS = [(4, 4), (4, 8), (8, 4), (8, 8), (4, 16), (16, 4), (8, 16), (16, 8), (16, 16), (8, 32), (32, 8), (16, 32), (32, 16), (32, 32),
          (16, 64), (64, 16), (32, 64), (64, 32), (64, 64), (64, 128), (128, 64), (128, 128)]


def pictures(w, h, seed, n_refs=3):
    """Source + references: the references are the source texture displaced by a fraction of a sample, plus noise."""
    rng = np.random.default_rng(seed)
    src = cm.Yuv(w, h, 8, pad=ic.REF_PAD)
    refs = []
    yy, xx = np.mgrid[0:h + 2 * ic.REF_PAD, 0:w + 2 * ic.REF_PAD].astype(np.float64)

    def tex(dx, dy):
        return 128 + 70 * np.sin((xx + dx) / 3.1) * np.cos((yy + dy) / 4.3) + 30 * np.sin((xx + dx + yy + dy) / 9.0)

    src.bufs[0][...] = np.clip(np.rint(tex(0, 0) + rng.integers(-4, 5, yy.shape)), 0, 255).astype(np.uint8)
    for i in range(n_refs):
        r = cm.Yuv(w, h, 8, pad=ic.REF_PAD)
        dx, dy = rng.uniform(-2.5, 2.5, 2)
        r.bufs[0][...] = np.clip(np.rint(tex(dx, dy) + rng.integers(-6, 7, yy.shape)), 0, 255).astype(np.uint8)
        refs.append(r)
    return src, refs


def limits(mi_rows, mi_cols, x, y, bw, bh, ref_row, ref_col):
    """svt_av1_set_mv_search_range (av1me.c:249-271) + svt_av1_set_subpel_mv_search_range (mcomp.h:124-138) on the UMV window
    of md_subpel_search (EbProductCodingLoop.c:2090-2096); checked against the reference in test_oracle_subpel.py."""
    mi_row, mi_col = y >> 2, x >> 2
    row_min, col_min = -(((mi_row + (bh >> 2)) * 4) + 4), -(((mi_col + (bw >> 2)) * 4) + 4)
    row_max, col_max = (mi_rows - mi_row) * 4 + 4, (mi_cols - mi_col) * 4 + 4
    MAXF, LOW, UPP = 1023, -(1 << 14), 1 << 14
    c0, r0 = (ref_col >> 3) - MAXF + int(ref_col & 7 != 0), (ref_row >> 3) - MAXF + int(ref_row & 7 != 0)
    c1, r1 = (ref_col >> 3) + MAXF, (ref_row >> 3) + MAXF
    c0, r0, c1, r1 = max(c0, (LOW >> 3) + 1), max(r0, (LOW >> 3) + 1), min(c1, (UPP >> 3) - 1), min(r1, (UPP >> 3) - 1)
    col_min, col_max, row_min, row_max = max(col_min, c0), min(col_max, c1), max(row_min, r0), min(row_max, r1)
    mx = MAXF * 8
    minc, maxc = max(col_min * 8, ref_col - mx), min(col_max * 8, ref_col + mx)
    minr, maxr = max(row_min * 8, ref_row - mx), min(row_max * 8, ref_row + mx)
    return max(LOW + 1, minc), min(UPP - 1, maxc), max(LOW + 1, minr), min(UPP - 1, maxr)


def make_jobs(w, h, n_refs, n_jobs, seed, blocks=BLOCKS):
    rng = np.random.default_rng(seed)
    mi_cols, mi_rows = 2 * ((w + 7) >> 3), 2 * ((h + 7) >> 3)
    jobs = np.zeros(n_jobs, sb.SUBPEL_JOB_DTYPE)
    for k in range(n_jobs):
        bw, bh = blocks[k % len(blocks)]
        if bw > w or bh > h:
            bw, bh = 8, 8
        x = int(rng.integers(0, (w - bw) // 4 + 1)) * 4
        y = int(rng.integers(0, (h - bh) // 4 + 1)) * 4
        if k % 9 == 0:  # at a picture corner, MV pointing outside: the limits cut the search
            x, y = (0 if k % 2 else w - bw), (0 if k % 4 < 2 else h - bh)
            fr, fc = ((-(y + bh + 4)) if k % 4 < 2 else (h - y + 4)), ((-(x + bw + 4)) if k % 2 else (w - x + 4))
        else:
            fr, fc = int(rng.integers(-12, 13)), int(rng.integers(-12, 13))
        j = jobs[k]
        j["blk_x"], j["blk_y"], j["bw"], j["bh"], j["ref"] = x, y, bw, bh, rng.integers(0, n_refs)
        j["start_mv_row"], j["start_mv_col"] = fr * 8, fc * 8
        j["ref_mv_row"], j["ref_mv_col"] = fr * 8 + int(rng.integers(-20, 21)), fc * 8 + int(rng.integers(-20, 21))
        j["col_min"], j["col_max"], j["row_min"], j["row_max"] = limits(mi_rows, mi_cols, x, y, bw, bh, int(j["ref_mv_row"]), int(j["ref_mv_col"]))
    return jobs


def cost_tables(seed):
    """nmv_vec_cost / nmvcoststack shaped like svt_av1_build_nmv_cost_table's output: a class term growing with log2|v| plus
    bit terms; returned as (mvjcost[4], [table0, table1]) with the tables centred at index MV_MAX."""
    rng = np.random.default_rng(seed)
    v = np.arange(-MV_MAX, MV_MAX + 1)
    tabs = []
    for c in range(2):
        mag = np.abs(v)
        cls = np.floor(np.log2(np.maximum(mag, 1) / 2.0 + 1)).astype(np.int64)
        cost = 300 + 512 * cls + 180 * (mag & 1) + 90 * ((mag >> 1) & 3) + rng.integers(0, 64, v.shape)
        cost[MV_MAX] = 120 + 10 * c
        tabs.append(np.ascontiguousarray(cost.astype(np.int32)))
    return [int(x) for x in rng.integers(300, 1800, 4)], tabs


def params(search_type=3, iters=2, allow_hp=1, forced_stop=0, cost_type=0, epb=120, seed=1):
    joint, tabs = cost_tables(seed)
    p = sb.SubpelParams(allow_hp, forced_stop, iters, search_type, cost_type, epb)
    for i in range(4):
        p.mvjcost[i] = joint[i]
    return p, tabs


def centre(t):
    return C.c_void_p(t.ctypes.data + 4 * MV_MAX)


def run_cpu(fn, p, tabs, src, refs, jobs):
    res = np.zeros(len(jobs), sb.SUBPEL_RESULT_DTYPE)
    arr = ic.frames_array(refs)
    ss = src.struct()
    rc = fn(C.byref(p), centre(tabs[0]), centre(tabs[1]), C.byref(ss), arr, len(refs), cm.ptr(np.ascontiguousarray(jobs)), len(jobs), cm.ptr(res))
    assert rc in (0, None) or fn.restype is None, rc
    return res
