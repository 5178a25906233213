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
