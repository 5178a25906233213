"""Shared test helpers: synthetic YUV, padded planes, ctypes access to the oracle (CPU restatement),
oracle/_ref (the real reference C path, when built) and the product library."""
import ctypes as C
import os
import subprocess

import numpy as np

import svtb200 as sb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libSvtAv1EncRef.so")
REFH_SO = os.path.join(ORACLE_DIR, "_ref", "librefharness.so")

_oracle = None
_ref = None
_refh = None


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"])
        _oracle = C.CDLL(ORACLE_SO)
        _oracle.orc_nxm_sad.restype = C.c_uint32
    return _oracle


def have_ref():
    return os.path.exists(REF_SO) and os.path.exists(REFH_SO)


def ref():
    """The unmodified reference C library (oracle/_ref), RTCD initialised to the C paths."""
    global _ref, _refh
    if _ref is None:
        _ref = C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)
        _refh = C.CDLL(REFH_SO)
        _refh.refh_init()
    return _ref


def refh():
    ref()
    return _refh


def ref_fn(name, restype=None):
    """Function behind an RTCD pointer variable of the reference (after setup_rtcd_internal(0))."""
    lib = ref()
    addr = C.c_void_p.in_dll(lib, name).value
    assert addr, name
    return C.CFUNCTYPE(restype)(addr) if False else C.cast(addr, C.CFUNCTYPE(restype))


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def synth_luma(width, height, n, seed=1234, noise=6):
    """SURVEY.md §8(d) generator: smooth moving texture + noise."""
    rng = np.random.default_rng(seed + 7919 * n)
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    v = 128 + 60 * np.sin((x + 3 * n) / 17.0) + 40 * np.cos((y - 2 * n) / 11.0)
    v = v + rng.integers(-noise, noise + 1, size=v.shape)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def pad_plane(img, geo):
    """Edge-replicated padded plane laid out as EbPictureBufferDesc describes (geo: sb.Plane)."""
    pad_x, pad_y = geo.origin_x, geo.origin_y
    h, w = img.shape
    assert (w, h) == (geo.width, geo.height) and geo.stride == w + 2 * pad_x
    return np.ascontiguousarray(np.pad(img, ((pad_y, pad_y), (pad_x, pad_x)), mode="edge"))


def me_planes(img, geos):
    """full / quarter / sixteenth padded planes of one picture (2x2 and 4x4 decimation, as the reference's
    downsample_decimation_input_picture does: keep the top-left sample)."""
    full = pad_plane(img, geos[0])
    q = pad_plane(np.ascontiguousarray(img[::2, ::2][: geos[1].height, : geos[1].width]), geos[1])
    s = pad_plane(np.ascontiguousarray(img[::4, ::4][: geos[2].height, : geos[2].width]), geos[2])
    return full, q, s


class MeBuffers:
    """Host result buffers of one picture's ME."""

    def __init__(self, n_sb):
        self.n_sb = n_sb
        self.best_sad = np.zeros((n_sb, 2, 4, 85), np.uint32)
        self.best_mv = np.zeros((n_sb, 2, 4, 85), np.uint32)
        self.hme = np.zeros((n_sb, 2, 4), np.dtype([("sc_x", "<i2"), ("sc_y", "<i2"), ("do_ref", "<u4"),
                                                     ("hme_sad", "<u8")]))
        self.me_mv = np.zeros((n_sb, 85 * 7, 2), np.int16)
        self.me_cand = np.zeros((n_sb, 85 * 23), np.uint8)
        self.total_cand = np.zeros((n_sb, 85), np.uint8)
        self.rc = np.zeros((n_sb,), np.uint32)

    def args(self):
        return [ptr(self.best_sad), ptr(self.best_mv), ptr(self.hme), ptr(self.me_mv), ptr(self.me_cand),
                ptr(self.total_cand), ptr(self.rc)]

    def fields(self):
        return dict(best_sad=self.best_sad, best_mv=self.best_mv, hme=self.hme, me_mv=self.me_mv,
                    me_cand=self.me_cand, total_cand=self.total_cand, rc=self.rc)


def make_me_case(width, height, n_l0, n_l1, seed=1234, motion=True):
    """Source picture + references (planes as numpy arrays) for an ME test."""
    geos = sb.me_geometry(width, height)
    src = me_planes(synth_luma(width, height, 8, seed), geos)
    refs = []
    for i in range(8):
        l, r = divmod(i, 4)
        n = 8 - (r + 1) if l == 0 else 8 + (r + 1)
        refs.append(me_planes(synth_luma(width, height, n if motion else 8, seed + (0 if motion else i)), geos))
    return geos, src, refs


def planes_struct(p3):
    return sb.MePlanes(p3[0].ctypes.data, p3[1].ctypes.data, p3[2].ctypes.data)


def run_oracle_me(params, src, refs):
    n_sb = ((params.full.width + 63) // 64) * ((params.full.height + 63) // 64)
    out = MeBuffers(n_sb)
    s = planes_struct(src)
    r = (sb.MePlanes * 8)(*[planes_struct(x) for x in refs])
    oracle().orc_me_picture(C.byref(params), C.byref(s), r, *out.args())
    return out


def run_ref_me(width, height, enc_mode, n_l0, n_l1, dist, temporal_layer, is_ref, geos, src, refs):
    n_sb = ((width + 63) // 64) * ((height + 63) // 64)
    out = MeBuffers(n_sb)
    params = sb.MeParams()
    s = planes_struct(src)
    r = (sb.MePlanes * 8)(*[planes_struct(x) for x in refs])
    d = (C.c_int32 * 8)(*[dist[i // 4][i % 4] for i in range(8)])
    rc = refh().refh_me_picture(width, height, enc_mode, n_l0, n_l1, d, temporal_layer, is_ref,
                                C.byref(geos[0]), C.byref(geos[1]), C.byref(geos[2]), C.byref(s), r,
                                C.byref(params), *out.args())
    assert rc == 0, rc
    return params, out


def assert_me_equal(a, b, params, what=""):
    """Compare two MeBuffers on every field the reference defines."""
    fa, fb = a.fields(), b.fields()
    for l in range(2):
        for r in range(4):
            used = l < params.num_lists and r < params.num_refs[l]
            if not used:
                continue
            np.testing.assert_array_equal(fa["hme"][:, l, r]["sc_x"], fb["hme"][:, l, r]["sc_x"], f"{what} sc_x {l},{r}")
            np.testing.assert_array_equal(fa["hme"][:, l, r]["sc_y"], fb["hme"][:, l, r]["sc_y"], f"{what} sc_y {l},{r}")
            np.testing.assert_array_equal(fa["hme"][:, l, r]["hme_sad"], fb["hme"][:, l, r]["hme_sad"], f"{what} hme_sad {l},{r}")
            np.testing.assert_array_equal(fa["hme"][:, l, r]["do_ref"], fb["hme"][:, l, r]["do_ref"], f"{what} do_ref {l},{r}")
            # best sad/mv are defined only where the integer search ran (do_ref after HME pruning): compare
            # where the SAD was initialised by the search (non-zero or both zero)
            np.testing.assert_array_equal(fa["best_sad"][:, l, r], fb["best_sad"][:, l, r], f"{what} best_sad {l},{r}")
            np.testing.assert_array_equal(fa["best_mv"][:, l, r], fb["best_mv"][:, l, r], f"{what} best_mv {l},{r}")
    for k in ("me_mv", "me_cand", "total_cand", "rc"):
        np.testing.assert_array_equal(fa[k], fb[k], f"{what} {k}")


# ------------------------------------------------------------------------------------------------------
# 4:2:0 frames for the EncDec / in-loop filter tests
# ------------------------------------------------------------------------------------------------------
class Yuv:
    """A 4:2:0 picture as three numpy planes (uint8 or uint16) with a little padding so strides != width."""

    def __init__(self, w, h, bit_depth=8, pad=16):
        self.w, self.h, self.bd = w, h, bit_depth
        dt = np.uint16 if bit_depth > 8 else np.uint8
        cw, ch = (w + 1) // 2, (h + 1) // 2
        self.pad = pad
        self.bufs = [np.zeros((h + 2 * pad, w + 2 * pad), dt), np.zeros((ch + 2 * pad, cw + 2 * pad), dt),
                     np.zeros((ch + 2 * pad, cw + 2 * pad), dt)]

    def plane(self, i):
        p = self.pad
        b = self.bufs[i]
        return b[p:b.shape[0] - p, p:b.shape[1] - p]

    def struct(self):
        p = self.pad
        ptrs = []
        for b in self.bufs:
            ptrs.append(b.ctypes.data + (p * b.shape[1] + p) * b.itemsize)
        return sb.Frame(ptrs[0], ptrs[1], ptrs[2], self.bufs[0].shape[1], self.bufs[1].shape[1], self.w, self.h, self.bd)

    def copy(self):
        o = Yuv(self.w, self.h, self.bd, self.pad)
        for a, b in zip(o.bufs, self.bufs):
            a[...] = b
        return o


def synth_yuv(w, h, n=0, seed=1234, bit_depth=8, noise=6):
    rng = np.random.default_rng(seed + 31 * n)
    f = Yuv(w, h, bit_depth)
    y = synth_luma(w, h, n, seed, noise).astype(np.int32)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    yy, xx = np.mgrid[0:ch, 0:cw].astype(np.float64)
    cb = np.clip(np.rint(128 + 50 * np.sin((xx + n) / 23.0) + rng.integers(-3, 4, (ch, cw))), 0, 255).astype(np.int32)
    cr = np.clip(np.rint(128 + 50 * np.cos((yy - n) / 19.0) + rng.integers(-3, 4, (ch, cw))), 0, 255).astype(np.int32)
    for i, p in enumerate((y, cb, cr)):
        if bit_depth > 8:
            p = (p << (bit_depth - 8)) + rng.integers(0, 1 << (bit_depth - 8), p.shape)
        f.plane(i)[...] = p
    return f


def degrade(f, seed=5, amp=10):
    """A 'reconstruction': the picture with coding-like noise + 8x8 blockiness."""
    rng = np.random.default_rng(seed)
    o = f.copy()
    mx = (1 << f.bd) - 1
    sc = 1 << (f.bd - 8)
    for i in range(3):
        p = o.plane(i).astype(np.int32)
        bs = 8 if i == 0 else 4
        blk = rng.integers(-amp, amp + 1, ((p.shape[0] + bs - 1) // bs, (p.shape[1] + bs - 1) // bs)) * sc
        p = p + np.kron(blk, np.ones((bs, bs), np.int32))[: p.shape[0], : p.shape[1]] + rng.integers(-3 * sc, 3 * sc + 1, p.shape)
        o.plane(i)[...] = np.clip(p, 0, mx)
    return o


def skip_map(mi_rows, mi_cols, seed=3, frac=0.3, all_skip_fb=True):
    rng = np.random.default_rng(seed)
    r8, c8 = (mi_rows + 1) // 2, (mi_cols + 1) // 2
    m = (rng.random((r8, c8 + 3)) < frac).astype(np.uint8)
    if all_skip_fb and r8 > 8 and c8 > 8:
        m[0:8, 8:16] = 1  # one filter block entirely skipped
    return m


# ------------------------------------------------------------------------------------------------------
# block partitions for the deblocking tests
# ------------------------------------------------------------------------------------------------------
# BlockSize enum values (Common/Codec/EbDefinitions.h:531-552)
_BS = {(4, 4): 0, (4, 8): 1, (8, 4): 2, (8, 8): 3, (8, 16): 4, (16, 8): 5, (16, 16): 6, (16, 32): 7, (32, 16): 8,
       (32, 32): 9, (32, 64): 10, (64, 32): 11, (64, 64): 12, (4, 16): 16, (16, 4): 17, (8, 32): 18, (32, 8): 19,
       (16, 64): 20, (64, 16): 21}


def random_partition(mi_rows, mi_cols, seed=0, p_split=0.55, p_inter=0.6, p_skip=0.4):
    """Per-mi arrays (sb_type, tx_depth, is_inter, skip) of a random AV1-like partition of the picture."""
    rng = np.random.default_rng(seed)
    sbt = np.zeros((mi_rows, mi_cols), np.uint8)
    dep = np.zeros((mi_rows, mi_cols), np.uint8)
    inter = np.zeros((mi_rows, mi_cols), np.uint8)
    skip = np.zeros((mi_rows, mi_cols), np.uint8)

    def leaf(r, c, h, w):
        hh, ww = min(h, mi_rows * 4 - r * 4), min(w, mi_cols * 4 - c * 4)
        if hh <= 0 or ww <= 0:
            return
        bs = _BS[(w, h)]
        d = int(rng.integers(0, 3))
        it, sk = int(rng.random() < p_inter), int(rng.random() < p_skip)
        r1, c1 = min(mi_rows, r + h // 4), min(mi_cols, c + w // 4)
        sbt[r:r1, c:c1], dep[r:r1, c:c1], inter[r:r1, c:c1], skip[r:r1, c:c1] = bs, d, it, sk

    def split(r, c, size):
        if r >= mi_rows or c >= mi_cols:
            return
        u = rng.random()
        if size > 4 and u < p_split:
            half = size // 2
            for dr in (0, half // 4):
                for dc in (0, half // 4):
                    split(r + dr, c + dc, half)
        elif size > 4 and u < p_split + 0.12:   # horizontal halves (w x h/2)
            leaf(r, c, size // 2, size)
            leaf(r + size // 8, c, size // 2, size)
        elif size > 4 and u < p_split + 0.24:   # vertical halves
            leaf(r, c, size, size // 2)
            leaf(r, c + size // 8, size, size // 2)
        elif size >= 16 and u < p_split + 0.30:  # 4:1 strips
            for k in range(4):
                leaf(r + k * size // 16, c, size // 4, size)
        else:
            leaf(r, c, size, size)

    for r in range(0, mi_rows, 16):
        for c in range(0, mi_cols, 16):
            split(r, c, 64)
    return sbt, dep, inter, skip
