"""Synthetic YUV 4:2:0 clips for the encoder-level runs (SURVEY.md §8(d) generator: smooth moving texture + noise so
ME finds non-zero MVs and residuals are non-trivial).  8-bit: planar uint8; 10-bit: planar little-endian uint16
(yuv420p10le: the 8-bit sample << 2 plus two random LSBs).  Frames are generated in chunks to bound memory.

    python tools/make_yuv.py out.yuv 1920 1080 300 [--bits 10] [--seed 1234]
"""
import argparse
import numpy as np


def frame_planes(w, h, n, seed=1234, bits=8, noise=6):
    rng = np.random.default_rng(seed + 7919 * n)
    # the texture is separable: a row term + a column term, broadcast (keeps 300 frames of 1080p to a few seconds)
    row = (128 + 60 * np.sin((np.arange(w, dtype=np.float64) + 3 * n) / 17.0))[None, :]
    col = (40 * np.cos((np.arange(h, dtype=np.float64) - 2 * n) / 11.0))[:, None]
    luma = np.clip(np.rint(row + col + rng.integers(-noise, noise + 1, size=(h, w))), 0, 255).astype(np.int32)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    cb_row = (128 + 50 * np.sin((np.arange(cw, dtype=np.float64) + n) / 23.0))[None, :]
    cr_col = (128 + 50 * np.cos((np.arange(ch, dtype=np.float64) - n) / 19.0))[:, None]
    cb = np.clip(np.rint(cb_row + rng.integers(-3, 4, (ch, cw))), 0, 255).astype(np.int32)
    cr = np.clip(np.rint(cr_col + rng.integers(-3, 4, (ch, cw))), 0, 255).astype(np.int32)
    out = []
    for p in (luma, cb, cr):
        if bits > 8:
            p = (p << (bits - 8)) + rng.integers(0, 1 << (bits - 8), p.shape)
            out.append(p.astype("<u2"))
        else:
            out.append(p.astype(np.uint8))
    return out


def write_clip(path, w, h, frames, bits=8, seed=1234):
    with open(path, "wb") as f:
        for n in range(frames):
            for p in frame_planes(w, h, n, seed, bits):
                f.write(p.tobytes())
    return path


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("out"); ap.add_argument("width", type=int); ap.add_argument("height", type=int)
    ap.add_argument("frames", type=int); ap.add_argument("--bits", type=int, default=8)
    ap.add_argument("--seed", type=int, default=1234)
    a = ap.parse_args()
    write_clip(a.out, a.width, a.height, a.frames, a.bits, a.seed)
