"""CPU-side checks of the encoder integration (no GPU needed): the overlay script still matches the reference tree, the
overlay build is inert without SVT_CUDA (bit-identical to the plain reference build), the API driver reproduces the
reference application's bitstream, the SIMD ("AVX2-minus-asm") reference build reproduces the C-only bitstream, the
CUDA backend fails loudly without a device, and bench.py's reference arm emits the contract's JSON line."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import encode_compare as ec  # noqa: E402
import make_yuv  # noqa: E402

B = os.path.join(ROOT, "integration", "_build")
built = all(os.path.exists(p) for p in list(ec.APPS.values()) + [os.path.join(B, "enc_bench_ref_simd"), os.path.join(B, "enc_bench_cuda_simd")])
need_build = pytest.mark.skipif(not built, reason="integration/_build or oracle/_ref/app not built (python __graft_entry__.py)")


@pytest.fixture(scope="module")
def clip(tmp_path_factory):
    d = tmp_path_factory.mktemp("enc")
    p = str(d / "clip_352x288.yuv")
    make_yuv.write_clip(p, 352, 288, 10, 8)
    return p, str(d)


@pytest.mark.skipif(not os.path.isdir("/root/reference/Source/Lib"), reason="reference tree not mounted")
def test_overlay_hooks_apply_to_the_reference_tree(tmp_path):
    """Every anchor of integration/overlay.py is found exactly once; the patch only ADDS lines inside `#if SVT_CUDA` or
    replaces one condition line."""
    subprocess.check_call([sys.executable, os.path.join(ROOT, "integration", "overlay.py")])
    patch = open(os.path.join(B, "overlay.patch")).read()
    assert patch.count("--- a/") == 7
    removed = [l for l in patch.splitlines() if l.startswith("-") and not l.startswith("---")]
    assert len(removed) == 1 and "loop_filter_mode == 1" in removed[0], removed
    for name in ("svt_cuda_backend_init", "svt_cuda_backend_deinit", "svt_cuda_me_segment", "svt_cuda_dlf_frame", "svt_cuda_cdef_picture", "svt_cuda_lr_frame", "svt_cuda_pa_statistics"):
        assert name in patch


@need_build
def test_overlay_build_is_inert_without_svt_cuda(clip):
    path, wd = clip
    runs = [ec.run_variant(v, path, 352, 288, 10, 8, 50, 8, wd) for v in ("ref_c", "ref_simd")]
    # cuda_* variants with the switch off: run_variant sets SVT_CUDA=1 for cuda builds, so call the apps with it unset
    for base in ("cuda_c", "cuda_simd"):
        apps = dict(ec.APPS)
        ec.APPS["tmp_" + base] = apps[base]
        try:
            runs.append(ec.run_variant("tmp_" + base, path, 352, 288, 10, 8, 50, 8, wd))
        finally:
            del ec.APPS["tmp_" + base]
    assert all(r["rc"] == 0 for r in runs), runs
    assert len({r["ivf_md5"] for r in runs}) == 1 and len({r["rec_md5"] for r in runs}) == 1, runs


@need_build
def test_api_driver_reproduces_the_application_bitstream(clip):
    path, wd = clip
    app = ec.run_variant("ref_simd", path, 352, 288, 10, 8, 50, 8, wd)
    out = os.path.join(wd, "drv.obu")
    r = subprocess.run([os.path.join(B, "enc_bench_ref_simd"), path, "352", "288", "10", "8", "8", "50", "0", out],
                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=600)
    assert r.returncode == 0
    res = json.loads([l for l in r.stdout.decode().splitlines() if l.startswith("{")][-1])
    assert res["packets"] == 10 and res["fps_all"] > 0
    sys.path.insert(0, ROOT)
    import bench
    assert ec.md5(out) == bench.ivf_payload_md5(os.path.join(wd, "ref_simd.ivf"))
    assert app["rc"] == 0


@need_build
def test_cuda_backend_without_a_device_is_an_error_not_a_fallback(clip):
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "svt-av1_b200", "libsvtav1_b200.so"))
    if lib.svt_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    path, wd = clip
    r = ec.run_variant("cuda_simd", path, 352, 288, 10, 8, 50, 8, wd, timeout=120)
    assert r["rc"] != 0 and r["ivf_md5"] is None
    assert any("no CPU fallback" in l for l in r.get("log", []) + r.get("tail", []))


@need_build
def test_bench_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=1200)
    assert r.returncode == 0
    lines = r.stdout.decode().splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "1080p30 8-bit preset-8 encode fps" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "reference-avx2-minus-asm" and d["cpu_baseline"]["cores"] == os.cpu_count()
    assert "configs[1]" in d["config"]["workload"]


@need_build
def test_gop_sharded_encode_splices_to_a_full_stream(tmp_path):
    """tools/shard_encode.py with the CPU reference encoder as the worker (no GPU): 3 GOPs over 2 ranks, every packet of the
    clip present once in the spliced stream."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "shard_encode.py"), "--width", "352", "--height", "288", "--frames", "40",
                        "--gop", "16", "--gpus", "2", "--sequential", "--variant", "ref_simd", "--preset", "8", "--qp", "50",
                        "--workdir", str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    d = json.loads(r.stdout.decode().strip().splitlines()[-1])
    assert d["packets"] == 40 and d["gops_per_rank"] == [2, 1]
