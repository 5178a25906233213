"""numpy restatement (TEST INFRASTRUCTURE) of the small reductions next to the transform chain and of the
single-position SAD family; pinned against the reference C functions in tests/test_oracle_misc.py.
Reference: Common/Codec/EbPictureOperators.c:156-231, Common/C_DEFAULT/EbPictureOperators_C.c:65-85,
Common/Codec/common_dsp_rtcd.c:47-69, Encoder/C_DEFAULT/EbComputeSAD_C.c:39-56, sad_av1.c."""
import numpy as np


def sad(a, b):
    return int(np.abs(a.astype(np.int64) - b.astype(np.int64)).sum())


def sse(a, b):
    d = a.astype(np.int64) - b.astype(np.int64)
    return int((d * d).sum())


def full_distortion32(coeff, recon):
    c, r = coeff.astype(np.int64), recon.astype(np.int64)
    return int(((c - r) ** 2).sum()), int((c * c).sum())


def satd(coeff):
    return int(np.abs(coeff.astype(np.int64)).sum())


def block_error(coeff, dq):
    # int (32-bit) products, 64-bit sums
    d = (coeff.astype(np.int64) - dq.astype(np.int64)).astype(np.int32)
    e = (d.astype(np.int64) * d.astype(np.int64)).astype(np.int32).astype(np.int64).sum()
    s = (coeff.astype(np.int64) * coeff.astype(np.int64)).astype(np.int32).astype(np.int64).sum()
    return int(e), int(s)


def subtract(src, pred):
    return (src.astype(np.int32) - pred.astype(np.int32)).astype(np.int16)


def cdef_dist(dst, dstride_blocks, src_blocks, dlist, bsize, coeff_shift, pli):
    """compute_cdef_dist_c / compute_cdef_dist_8bit_c (Encoder/Codec/EbEncCdef.c:134-220).  dst: 2-D array (the filter
    block), src_blocks: [count][bh*bw] packed source blocks, dlist: [(by, bx)], bsize: BlockSize 0..3."""
    import math
    bw = 8 if bsize in (2, 3) else 4
    bh = 8 if bsize in (1, 3) else 4
    total = 0
    for bi, (by, bx) in enumerate(dlist):
        d = dst[by * bh:(by + 1) * bh, bx * bw:(bx + 1) * bw].astype(np.int64).reshape(-1)
        s = src_blocks[bi].astype(np.int64)
        if bsize == 3 and pli == 0:  # dist_8x8 (:75-98 / :20-33): double arithmetic in the reference's order
            sum_s, sum_d = int(s.sum()), int(d.sum())
            sum_s2, sum_d2, sum_sd = int((s * s).sum()), int((d * d).sum()), int((s * d).sum())
            svar = sum_s2 - ((sum_s * sum_s + 32) >> 6)
            dvar = sum_d2 - ((sum_d * sum_d + 32) >> 6)
            a = float(sum_d2 + sum_s2 - 2 * sum_sd) * .5
            b = a * float(svar + dvar + (400 << 2 * coeff_shift))
            c = math.sqrt(float(20000 << 4 * coeff_shift) + float(svar) * float(dvar))
            total += int(math.floor(.5 + b / c))
        else:
            total += int(((d - s) ** 2).sum())
    return total >> (2 * coeff_shift)


def compute_stats(win, dgd, src, h_start, h_end, v_start, v_end, bit_depth=8):
    """svt_av1_compute_stats_c / _highbd_c (Encoder/Codec/EbRestorationPick.c:704-790).  dgd, src: 2-D integer arrays
    indexed [row, col] with enough margin around the unit; returns (M[win^2], H[win^2, win^2]) as int64."""
    hw = win // 2
    d = dgd.astype(np.int64)
    unit = d[v_start:v_end, h_start:h_end]
    avg = int(unit.sum()) // unit.size  # find_average: truncating division, stored in the sample type
    x = (src.astype(np.int64)[v_start:v_end, h_start:h_end] - avg).reshape(-1)
    cols = []
    for k in range(-hw, hw + 1):          # column offset (outer), row offset (inner): idx = (k+hw)*win + (l+hw)
        for l in range(-hw, hw + 1):
            cols.append((d[v_start + l:v_end + l, h_start + k:h_end + k] - avg).reshape(-1))
    Y = np.stack(cols, axis=1)
    M = Y.T @ x
    H = Y.T @ Y
    if bit_depth > 8:
        div = 16 if bit_depth == 12 else 4
        M = np.fix(M / div).astype(np.int64) if False else (np.sign(M) * (np.abs(M) // div))  # C division truncates
        H = np.sign(H) * (np.abs(H) // div)
    return M.astype(np.int64), H.astype(np.int64)


def pixel_proj_error(src, dat, flt0, flt1, xq, r, highbd):
    """svt_av1_lowbd_pixel_proj_error_c / svt_av1_highbd_pixel_proj_error_c (EbRestorationPick.c:174-315)."""
    s, d = src.astype(np.int64), dat.astype(np.int64)
    if r[0] > 0 or r[1] > 0:
        u = d << 4
        v = np.full_like(u, 1 << 10) if highbd else (u << 7)
        if r[0] > 0:
            v = v + xq[0] * (flt0.astype(np.int64) - u)
        if r[1] > 0:
            v = v + xq[1] * (flt1.astype(np.int64) - u)
        e = (v >> 11) + d - s if highbd else ((v + (1 << 10)) >> 11) - s
    else:
        e = d - s
    return int((e * e).sum())


def get_proj_subspace(src, dat, flt0, flt1, r):
    """svt_get_proj_subspace_c (EbRestorationPick.c:337-440): xq[2].  Sums as exact integers, solve in float64 with the
    reference's expressions (Python floats are IEEE doubles; no contraction)."""
    import math
    u = dat.astype(np.int64) << 4
    s = (src.astype(np.int64) << 4) - u
    f1 = (flt0.astype(np.int64) - u) if r[0] > 0 else np.zeros_like(u)
    f2 = (flt1.astype(np.int64) - u) if r[1] > 0 else np.zeros_like(u)
    size = u.size
    h00, h11, h01 = float(int((f1 * f1).sum())) / size, float(int((f2 * f2).sum())) / size, float(int((f1 * f2).sum())) / size
    c0, c1 = float(int((f1 * s).sum())) / size, float(int((f2 * s).sum())) / size

    def rint(v):  # round half to even, as rint() in the default rounding mode
        return int(round(v))
    if r[0] == 0:
        return [0, 0] if h11 < 1e-8 else [0, rint(c1 / h11 * 128)]
    if r[1] == 0:
        return [0, 0] if h00 < 1e-8 else [rint(c0 / h00 * 128), 0]
    det = h00 * h11 - h01 * h01
    if det < 1e-8:
        return [0, 0]
    return [rint((h11 * c0 - h01 * c1) / det * 128), rint((h00 * c1 - h01 * c0) / det * 128)]


def variance(a, b):
    """svt_aom_varianceWxH_c / svt_aom_mse16x16_c (EbComputeVariance_C.c:14-61, EbPsnr.c:84): (variance, sse), 32-bit wrap."""
    d = a.astype(np.int64) - b.astype(np.int64)
    sse = int((d * d).sum()) & 0xFFFFFFFF
    s = int(d.sum())
    q = abs(s * s) // a.size  # C division of a non-negative value
    return (sse - (q & 0xFFFFFFFF)) & 0xFFFFFFFF, sse
