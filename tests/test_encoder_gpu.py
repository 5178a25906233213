"""Encoder-level parity: the reference encoder with the CUDA backend bound in
(integration/_build, SVT_CUDA=1) must produce the SAME .ivf bitstream and the SAME reconstruction as the plain C-only
reference encoder on the same synthetic YUV, through the reference's own application and API.  The encoder is
deterministic, so the comparison is an md5 equality.  Everything here runs the prebuilt binaries (the GPU box has no
/root/reference)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import encode_compare as ec  # noqa: E402
import make_yuv  # noqa: E402

pytestmark = pytest.mark.gpu

need = [ec.APPS["ref_c"], ec.APPS["cuda_c"], ec.APPS["ref_simd"], ec.APPS["cuda_simd"]]
have = all(os.path.exists(p) for p in need)
skip_if_unbuilt = pytest.mark.skipif(not have, reason="integration/_build or oracle/_ref/app not built (python __graft_entry__.py)")


def _clip(tmp_path, w, h, frames, bits):
    p = str(tmp_path / ("clip_%dx%d_%d_%d.yuv" % (w, h, frames, bits)))
    make_yuv.write_clip(p, w, h, frames, bits)
    return p


def _same(a, b):
    assert a["rc"] == 0 and b["rc"] == 0, (a, b)
    assert a["ivf_md5"] and a["ivf_md5"] == b["ivf_md5"], (a, b)
    assert a["rec_md5"] and a["rec_md5"] == b["rec_md5"], (a, b)


@skip_if_unbuilt
@pytest.mark.parametrize("stages", ["me", "dlf", "cdef", "me+dlf+cdef"])
def test_config1_360p_preset8_bitstream_and_recon_md5(tmp_path, stages):
    """BASELINE configs[0]: 640x360 8-bit, 30 frames, preset 8, CQP qp 50 - each GPU stage alone and all together."""
    clip = _clip(tmp_path, 640, 360, 30, 8)
    ref = ec.run_variant("ref_c", clip, 640, 360, 30, 8, 50, 8, str(tmp_path))
    gpu = ec.run_variant("cuda_c:" + stages, clip, 640, 360, 30, 8, 50, 8, str(tmp_path), extra_env={"SVT_CUDA_PROFILE": "1"})
    _same(ref, gpu)
    # the stage really ran on the GPU (the backend's profile line counts engine calls)
    log = " ".join(gpu.get("log", []))
    assert "ME pictures" in log
    import re
    m = re.search(r"engine: (\d+) ME pictures .*?, (\d+) dlf, (\d+) cdef", log)
    n_me, n_dlf, n_cdef = (int(x) for x in m.groups())
    assert (n_me > 0) == ("me" in stages) and (n_dlf > 0) == ("dlf" in stages) and (n_cdef > 0) == ("cdef" in stages), log


@skip_if_unbuilt
def test_360p_preset8_device_side_me_downsample(tmp_path):
    """SVT_CUDA_ME_DS=1: only the full-resolution luma is uploaded, the 1/4 and 1/16 HME planes are derived on the device
    (svt_b200_me_downsample) - the encode must not change."""
    clip = _clip(tmp_path, 640, 360, 20, 8)
    ref = ec.run_variant("ref_c", clip, 640, 360, 20, 8, 50, 8, str(tmp_path))
    gpu = ec.run_variant("cuda_c:me", clip, 640, 360, 20, 8, 50, 8, str(tmp_path), extra_env={"SVT_CUDA_ME_DS": "1"})
    _same(ref, gpu)


@skip_if_unbuilt
def test_10bit_preset6_bitstream_and_recon_md5(tmp_path):
    """A 10-bit preset-6 clip (the configs[2] stage set at a small size): 16-bit pipeline, loop_filter_mode 3 (the frame is
    deblocked in dlf_kernel), CDEF on 16-bit planes, restoration on the CPU."""
    clip = _clip(tmp_path, 640, 360, 12, 10)
    ref = ec.run_variant("ref_c", clip, 640, 360, 12, 6, 50, 10, str(tmp_path))
    gpu = ec.run_variant("cuda_c", clip, 640, 360, 12, 6, 50, 10, str(tmp_path))
    _same(ref, gpu)


@skip_if_unbuilt
def test_simd_build_with_cuda_backend_matches_simd_reference(tmp_path):
    """The speed build (reference AVX2/AVX-512 code + CUDA backend) against the reference's SIMD encoder, ragged size
    (not a multiple of 64), through the API driver the bench uses as well."""
    clip = _clip(tmp_path, 720, 400, 24, 8)
    ref = ec.run_variant("ref_simd", clip, 720, 400, 24, 8, 43, 8, str(tmp_path))
    gpu = ec.run_variant("cuda_simd", clip, 720, 400, 24, 8, 43, 8, str(tmp_path))
    _same(ref, gpu)
