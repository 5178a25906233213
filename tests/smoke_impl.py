"""smoke(): one small invocation of every stage of the hot path on cuda:0 — open-loop ME, the fused EncDec transform-unit
kernel, deblocking, the CDEF search and apply — and of the rows widened into (sub-pel search, inter prediction, HME plane
downsampling), each checked bit-exactly against the CPU oracle through the same code the `-m gpu` tests use."""
import common as cm
import svtb200 as sb


def run():
    import torch
    assert torch.cuda.is_available(), "smoke() needs a GPU"
    import gpu_runner as gr
    lib = sb.load()
    W, H = 256, 192
    dist = ((1, 2, 3, 4), (1, 2, 3, 4))
    params = sb.preset8_me_params(W, H, 2, 1, dist, 1, 1)
    geos, src, refs = cm.make_me_case(W, H, 2, 1)
    want = cm.run_oracle_me(params, src, refs)
    got = gr.run_gpu_me(params, src, refs)
    cm.assert_me_equal(got, want, params, "smoke gpu-vs-oracle")
    done = ["me"]
    import test_cdef_gpu
    import test_dlf_gpu
    import test_interp_gpu
    import test_me_gpu
    import test_subpel_gpu
    import test_txfm_gpu
    test_txfm_gpu.test_encode_tus_vs_oracle((2, 8, 0))
    test_txfm_gpu.test_encode_tus_vs_oracle((3, 10, 1))
    done.append("encdec")
    test_dlf_gpu.test_dlf_frame_vs_oracle(test_dlf_gpu.DLF_CASES[0])
    done.append("dlf")
    test_cdef_gpu.test_cdef_search_vs_oracle((192, 136, 8, 3))
    test_cdef_gpu.test_cdef_apply_vs_oracle((192, 136, 8))
    done.append("cdef")
    test_subpel_gpu.test_subpel_search_vs_oracle(0)
    test_interp_gpu.test_inter_predict_vs_oracle(176, 144, 8, 64, 21, "texture")
    test_me_gpu.test_me_downsample_vs_oracle(176, 144, 1)
    done += ["subpel", "inter_predict", "me_downsample"]
    # round 2: the tensor-core Wiener statistics, the temporal filter, picture statistics, open-loop intra search and the CDEF
    # strength decision - the -m gpu tests themselves, one small case each
    import test_misc_gpu
    import test_ois_gpu
    import test_pa_gpu
    import test_tf_gpu
    test_misc_gpu.test_compute_stats_extreme_values(8)
    done.append("wiener_stats(imma)")
    import tf_cases
    test_tf_gpu.test_planewise_block_dropin_vs_oracle(tf_cases.CASES[0])
    done.append("temporal_filter")
    test_pa_gpu.test_picture_mean_variance_vs_oracle((200, 120, 0))
    test_ois_gpu.test_ois_dc_picture_vs_oracle((200, 120))
    done.append("picture_stats, open_loop_intra")
    test_cdef_gpu.test_cdef_decide_vs_oracle((0, 3, 68, 120, "smooth"))
    done.append("cdef_decide")
    print("smoke OK: %s bit-exact vs oracle, launches=%d" % (", ".join(done), lib.svt_b200_launch_count()))
