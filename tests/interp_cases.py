"""Inter-prediction test cases: random block partitions of a picture turned into the per-plane jobs av1_inter_prediction
(EbEncInterPrediction.c:4040-4930) issues, reference pictures with a replicated border, and the oracle / reference runners."""
import ctypes as C

import numpy as np

import common as cm
import svtb200 as sb

REF_PAD = 160  # >= 128 + 4 + 4 + 7: how far clamp_mv_to_umv_border_sb lets the largest block + its taps reach
JNT_WEIGHTS = [(9, 7), (11, 5), (12, 4), (13, 3), (7, 9), (5, 11), (4, 12), (3, 13)]  # quant_dist_lookup_table (:304-307)


def ref_picture(w, h, bd, seed, kind="texture"):
    """A reconstructed-looking picture with EbPictureBufferDesc-style replicated padding."""
    rng = np.random.default_rng(seed)
    f = cm.Yuv(w, h, bd, pad=REF_PAD)
    mx = (1 << bd) - 1
    for i in range(3):
        p = f.plane(i)
        if kind == "rand":
            v = rng.integers(0, mx + 1, p.shape)
        elif kind == "extreme":
            v = rng.choice([0, mx], p.shape)
        else:
            yy, xx = np.mgrid[0:p.shape[0], 0:p.shape[1]]
            v = (mx / 2) * (1 + 0.45 * np.sin(xx / (5.0 + i)) * np.cos(yy / (7.0 - i))) + rng.integers(-12, 13, p.shape) * (mx // 255)
        p[...] = np.clip(np.rint(v), 0, mx).astype(p.dtype)
        f.bufs[i][...] = np.pad(p, REF_PAD, mode="edge")
    return f


def _leaves(x, y, n, rng, out, min_n=4, square_only=False):
    """Random AV1 partition of the n x n square at (x, y): split, or one of none / horz / vert / horz4 / vert4."""
    if n > min_n and rng.random() < (0.9 if n == 64 else 0.55 if n == 32 else 0.4 if n == 16 else 0.3):
        for dy in (0, n // 2):
            for dx in (0, n // 2):
                _leaves(x + dx, y + dy, n // 2, rng, out, min_n, square_only)
        return
    kind = rng.integers(0, 5) if n >= 8 and not square_only else 0
    if kind == 0:
        out.append((x, y, n, n))
    elif kind == 1:
        out += [(x, y, n, n // 2), (x, y + n // 2, n, n // 2)]
    elif kind == 2:
        out += [(x, y, n // 2, n), (x + n // 2, y, n // 2, n)]
    elif kind == 3 and n >= 16:
        out += [(x, y + k * n // 4, n, n // 4) for k in range(4)]
    elif kind == 4 and n >= 16:
        out += [(x + k * n // 4, y, n // 4, n) for k in range(4)]
    else:
        out.append((x, y, n, n))


def make_jobs(w, h, n_ref_frames, seed, sb_size=64, mv_range=96, far_mv_every=37, compound_frac=0.35, min_n=4, square_only=False):
    """Jobs of a whole picture as a numpy record array (sb.INTER_JOB_DTYPE). min_n=8, square_only=True: the block sizes preset 8
    itself uses (square blocks down to 8x8, no 4xN / Nx4)."""
    rng = np.random.default_rng(seed)
    mi_cols, mi_rows = 2 * ((w + 7) >> 3), 2 * ((h + 7) >> 3)
    blocks = []
    for sy in range(0, h, sb_size):
        for sx in range(0, w, sb_size):
            if sb_size == 128 and rng.random() < 0.3:
                blocks.append((sx, sy, 128, 128))
                continue
            for oy in range(0, sb_size, 64):
                for ox in range(0, sb_size, 64):
                    _leaves(sx + ox, sy + oy, 64, rng, blocks, min_n, square_only)
    jobs = []
    last_mode = {}  # (x>>2, y>>2) -> (mv_row, mv_col, ref) of the block covering that mi: the sub8x8 chroma case reads it

    def new_mode(k):
        far = far_mv_every and k % far_mv_every == 0
        rngv = 1400 if far else mv_range
        n_refs = 2 if (n_ref_frames > 1 and rng.random() < compound_frac) else 1
        refs = rng.choice(n_ref_frames, 2, replace=n_ref_frames < 2)
        mv = rng.integers(-rngv, rngv + 1, (2, 2))
        if k % 5 == 0:
            mv[:, 0] &= ~7  # whole-sample rows -> the x-only forms
        if k % 7 == 0:
            mv[:, 1] &= ~7
        if k % 11 == 0:
            mv[...] &= ~7  # the copy forms (luma; chroma still has a half-sample phase)
        if k % 13 == 0:
            mv[...] &= ~15  # copy in chroma too
        wgt = JNT_WEIGHTS[rng.integers(0, 8)]
        return dict(n_refs=n_refs, ref=refs, mv=mv, fx=int(rng.integers(0, 4)), fy=int(rng.integers(0, 4)),
                    jnt=int(rng.integers(0, 2)), fwd=wgt[0], bck=wgt[1])

    def job(plane, m, bw, bh, dx, dy, px, py, edges, single_ref=None):
        j = np.zeros((), sb.INTER_JOB_DTYPE)
        j["plane"], j["bw"], j["bh"] = plane, bw, bh
        j["n_refs"] = 1 if single_ref is not None else m["n_refs"]
        j["ref"] = m["ref"] if single_ref is None else [single_ref[2], 0]
        j["filter_x"], j["filter_y"] = m["fx"], m["fy"]
        j["use_jnt_comp_avg"], j["fwd_offset"], j["bck_offset"] = m["jnt"], m["fwd"], m["bck"]
        j["dst_x"], j["dst_y"], j["pre_x"], j["pre_y"] = dx, dy, px, py
        if single_ref is None:
            j["mv_row"], j["mv_col"] = m["mv"][:, 0], m["mv"][:, 1]
        else:
            j["mv_row"][0], j["mv_col"][0] = single_ref[0], single_ref[1]
        j["mb_to_left_edge"], j["mb_to_right_edge"], j["mb_to_top_edge"], j["mb_to_bottom_edge"] = edges
        jobs.append(j)

    for k, (x, y, bw, bh) in enumerate(blocks):
        if x >= w or y >= h:
            continue
        m = new_mode(k)
        edges = (-(x * 8), (mi_cols * 4 - bw - x) * 8, -(y * 8), (mi_rows * 4 - bh - y) * 8)
        job(0, m, bw, bh, x, y, x, y, edges)
        for my in range(y >> 2, (y + bh) >> 2):
            for mx_ in range(x >> 2, (x + bw) >> 2):
                last_mode[(mx_, my)] = (int(m["mv"][0, 0]), int(m["mv"][0, 1]), int(m["ref"][0]))
        has_uv = (bw > 4 or (x & 4)) and (bh > 4 or (y & 4))  # the last block of an 8x8 carries its chroma
        if not has_uv:
            continue
        cx, cy = ((x >> 3) << 3) // 2, ((y >> 3) << 3) // 2
        if (bw == 4 or bh == 4) and m["n_refs"] == 1 and rng.random() < 0.6:
            # sub8x8_inter (:4150-4330): one b4 job per covered luma block, each with that block's MV and reference
            b4w, b4h = bw >> 1, bh >> 1
            b8w, b8h = max(bw, 8) >> 1, max(bh, 8) >> 1
            row0, col0 = (-1 if bh == 4 else 0), (-1 if bw == 4 else 0)
            row = row0
            for yy in range(0, b8h, b4h):
                col = col0
                for xx in range(0, b8w, b4w):
                    nb = last_mode[((x >> 2) + col, (y >> 2) + row)]
                    for plane in (1, 2):
                        job(plane, m, b4w, b4h, cx + xx, cy + yy, cx + xx, cy + yy, edges, single_ref=nb)
                    col += 1
                row += 1
        else:
            for plane in (1, 2):
                job(plane, m, max(bw >> 1, 4), max(bh >> 1, 4), cx, cy, cx, cy, edges)
    return np.array(jobs, dtype=sb.INTER_JOB_DTYPE)


def frames_array(frames):
    return (sb.Frame * len(frames))(*[f.struct() for f in frames])


def run_cpu(fn, refs, pred, jobs):
    """fn = oracle().orc_inter_predict or refh().refh_inter_predict."""
    arr = frames_array(refs)
    ps = pred.struct()
    jobs = np.ascontiguousarray(jobs)
    fn(arr, len(refs), C.byref(ps), cm.ptr(jobs), len(jobs))
    return pred


def conv_rounds(bd, compound):
    """get_conv_params_no_round (convolve.h:44-71)."""
    r0 = 3
    r1 = 7 if compound else 14 - r0
    rng_ = bd + 7 - r0 + 2
    if rng_ > 16:
        r0 += rng_ - 16
        if not compound:
            r1 -= rng_ - 16
    return r0, r1
