"""svt_b200_ois_dc_picture (open-loop intra search of the TPL path, presets >= 5) against the oracle, which
tests/test_oracle_ois.py pins to the reference's open_loop_intra_search_mb; the oracle reads an edge-replicated padded
plane as the reference does, the GPU entry clamps its reads to the picture."""
import ctypes as C

import numpy as np
import pytest

import common as cm
import svtb200 as sb

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("geom", [(1920, 1080), (640, 360), (200, 120), (72, 88), (3840, 2160), (16, 16)])
def test_ois_dc_picture_vs_oracle(geom):
    import torch
    import gpu_runner as gr
    w, h = geom
    lib, orc = sb.load(), cm.oracle()
    pic = cm.synth_yuv(w, h, 3, 91, 8, noise=12)
    mbw, mbh = (w + 15) // 16, (h + 15) // 16
    pad = np.ascontiguousarray(np.pad(pic.plane(0), ((0, mbh * 16 - h + 16), (0, mbw * 16 - w + 16)), mode="edge"))
    want = np.zeros(mbw * mbh, np.int64)
    orc.orc_ois_dc_picture(cm.ptr(pad), pad.shape[1], w, h, cm.ptr(want))
    d = gr.DevYuv(pic)
    got = torch.full((mbw * mbh,), -5, dtype=torch.int64, device="cuda")
    st = d.struct()
    sb.check(lib.svt_b200_ois_dc_picture(C.byref(st), got.data_ptr(), None), lib)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # host-pointer form (the encoder binding's entry)
    host = np.zeros(mbw * mbh, np.int64)
    y = np.ascontiguousarray(pic.plane(0))
    sb.check(lib.svt_b200_ois_dc_picture_host(cm.ptr(y), y.shape[1], w, h, cm.ptr(host)), lib)
    np.testing.assert_array_equal(host, want)
    assert want.max() > 0
