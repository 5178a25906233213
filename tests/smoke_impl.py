"""smoke(): one small ME picture on cuda:0 vs the CPU oracle (bit-exact)."""
import common as cm
import svtb200 as sb


def run():
    import torch
    assert torch.cuda.is_available(), "smoke() needs a GPU"
    import gpu_runner as gr
    W, H = 256, 192
    dist = ((1, 2, 3, 4), (1, 2, 3, 4))
    params = sb.preset8_me_params(W, H, 2, 1, dist, 1, 1)
    geos, src, refs = cm.make_me_case(W, H, 2, 1)
    want = cm.run_oracle_me(params, src, refs)
    got = gr.run_gpu_me(params, src, refs)
    cm.assert_me_equal(got, want, params, "smoke gpu-vs-oracle")
    print("smoke OK: ME picture %dx%d bit-exact vs oracle, launches=%d" % (W, H, sb.load().svt_b200_launch_count()))
