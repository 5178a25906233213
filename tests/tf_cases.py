"""Shared case generator for the temporal-filter tests (oracle vs reference on the CPU, CUDA vs oracle on the GPU)."""
import ctypes as C

import numpy as np


class TfCase:
    pass


def make_case(seed, bd, chroma=1, bw=32, bh=32, split=None, big_motion=False, noise=(2.5, 1.2, 0.9), decay=4, amp=12):
    rng = np.random.default_rng(seed)
    c = TfCase()
    c.bd, c.chroma, c.bw, c.bh, c.decay = bd, chroma, bw, bh, decay
    dt = np.uint16 if bd > 8 else np.uint8
    top = (1 << bd) - 1
    yy, xx = np.mgrid[0:bh, 0:bw]
    base = ((np.sin(xx / 5.0 + seed) + np.cos(yy / 4.0) + 2) * top / 4)
    c.ys, c.ps = 40 + (seed % 3), 64  # strides: source picture row pitch, prediction block pitch (BW)
    c.uvs, c.ups = 24 + (seed % 5), 32
    def plane(h, w, stride, arr):
        buf = np.zeros((h, stride), dt)
        buf[:, :w] = np.clip(arr, 0, top).astype(dt)
        return buf
    scale = amp << (bd - 8)
    c.y_src = plane(bh, bw, c.ys, base + rng.integers(-3, 4, (bh, bw)))
    c.y_pre = plane(bh, bw, c.ps, base + rng.integers(-scale, scale + 1, (bh, bw)) * (rng.random((bh, bw)) < 0.5))
    cb = base[::2, ::2] * 0.5 + top / 4
    c.u_src = plane(bh // 2, bw // 2, c.uvs, cb + rng.integers(-2, 3, (bh // 2, bw // 2)))
    c.v_src = plane(bh // 2, bw // 2, c.uvs, top - cb + rng.integers(-2, 3, (bh // 2, bw // 2)))
    c.u_pre = plane(bh // 2, bw // 2, c.ups, cb + rng.integers(-scale // 2, scale // 2 + 1, (bh // 2, bw // 2)))
    c.v_pre = plane(bh // 2, bw // 2, c.ups, top - cb + rng.integers(-scale // 2, scale // 2 + 1, (bh // 2, bw // 2)))
    c.block_row, c.block_col = int(rng.integers(0, 2)), int(rng.integers(0, 2))
    c.split = np.array([int(rng.integers(0, 2)) for _ in range(4)] if split is None else [split] * 4, np.int32)
    hb = 16 if bd > 8 else 1
    c.err16 = rng.integers(0, 256 * 40 * hb, 16).astype(np.uint64)
    c.err32 = rng.integers(0, 1024 * 40 * hb, 4).astype(np.uint64)
    m = 400 if big_motion else 12
    c.mvx16, c.mvy16 = rng.integers(-m, m + 1, 16).astype(np.int16), rng.integers(-m, m + 1, 16).astype(np.int16)
    c.mvx32, c.mvy32 = rng.integers(-m, m + 1, 4).astype(np.int16), rng.integers(-m, m + 1, 4).astype(np.int16)
    c.min_frame_size = int(rng.choice([288, 360, 1080, 7]))
    c.noise = (C.c_double * 3)(*noise)
    # accumulators start non-zero: the filter ADDS into them
    c.y_acc0 = rng.integers(0, 1 << 20, (bh, c.ps)).astype(np.uint32)
    c.y_cnt0 = rng.integers(0, 3000, (bh, c.ps)).astype(np.uint16)
    c.u_acc0 = rng.integers(0, 1 << 20, (bh // 2, c.ups)).astype(np.uint32)
    c.u_cnt0 = rng.integers(0, 3000, (bh // 2, c.ups)).astype(np.uint16)
    c.v_acc0 = rng.integers(0, 1 << 20, (bh // 2, c.ups)).astype(np.uint32)
    c.v_cnt0 = rng.integers(0, 3000, (bh // 2, c.ups)).astype(np.uint16)
    return c


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def fresh(c):
    return [a.copy() for a in (c.y_acc0, c.y_cnt0, c.u_acc0, c.u_cnt0, c.v_acc0, c.v_cnt0)]


def run_reference(refh, c, via_rtcd=0):
    out = fresh(c)
    rc = refh.refh_tf_planewise(via_rtcd, c.bd, c.chroma, c.block_row, c.block_col, ptr(c.split), ptr(c.err16), ptr(c.err32), ptr(c.mvx16),
                                ptr(c.mvy16), ptr(c.mvx32), ptr(c.mvy32), c.min_frame_size, ptr(c.y_src), c.ys, ptr(c.y_pre), c.ps,
                                ptr(c.u_src), ptr(c.v_src), c.uvs, ptr(c.u_pre), ptr(c.v_pre), c.ups, c.bw, c.bh, 1, 1, c.noise, c.decay,
                                *[ptr(a) for a in out])
    assert rc == 0
    return out


def factors(orc, c):
    """den[3], block_error[4], d_factor[4] of the block (the host glue of the drop-in, restated in the oracle)."""
    den, be, df = (C.c_double * 3)(), (C.c_double * 4)(), (C.c_double * 4)()
    orc.orc_tf_den(c.decay, c.noise, den)
    i32 = c.block_col + c.block_row * 2
    orc.orc_tf_block_factors(int(c.split[i32]), ptr(c.err16[i32 * 4:i32 * 4 + 4].copy()), C.c_uint64(int(c.err32[i32])),
                             ptr(c.mvx16[i32 * 4:i32 * 4 + 4].copy()), ptr(c.mvy16[i32 * 4:i32 * 4 + 4].copy()),
                             C.c_int16(int(c.mvx32[i32])), C.c_int16(int(c.mvy32[i32])), c.min_frame_size, 1 if c.bd > 8 else 0, be, df)
    return den, be, df


def run_oracle(orc, c):
    out = fresh(c)
    den, be, df = factors(orc, c)
    orc.orc_tf_planewise(ptr(c.y_src), c.ys, ptr(c.y_pre), c.ps, ptr(c.u_src), ptr(c.v_src), c.uvs, ptr(c.u_pre), ptr(c.v_pre), c.ups,
                         c.bw, c.bh, 1, 1, den, be, df, c.chroma, c.bd, *[ptr(a) for a in out])
    return out


CASES = [dict(seed=s, bd=bd, chroma=ch, split=sp, big_motion=bm, noise=nz, decay=dc, amp=amp)
         for s, (bd, ch, sp, bm, nz, dc, amp) in enumerate([
             (8, 1, None, False, (2.5, 1.2, 0.9), 4, 12), (8, 1, 1, True, (0.3, 0.2, 0.1), 3, 30), (8, 0, 0, False, (6.0, 3.0, 2.0), 2, 4),
             (8, 1, 0, True, (0.0, 0.0, 0.0), 4, 60), (10, 1, None, False, (2.5, 1.2, 0.9), 4, 12), (10, 1, 1, True, (0.7, 0.4, 0.3), 3, 40),
             (10, 0, 0, False, (4.0, 2.0, 1.0), 4, 6), (8, 1, None, False, (1.0, 1.0, 1.0), 4, 1), (10, 1, None, True, (9.0, 9.0, 9.0), 2, 100)])]
