"""One stream encoded by several encoder instances, one per GPU, sharded by CLOSED GOP (BASELINE
configs[4]): GOP g (keyint frames, opening with an IDR: --irefresh-type 2 is the encoder's default) goes to rank g mod N;
rank r encodes its GOPs as one sub-stream on GPU r; the sub-streams' packets are spliced back GOP by GOP.  Nothing crosses
an IDR in the bitstream, so the splice is a valid AV1 stream; what the single-instance encoder shares across GOPs is
encoder-side look-ahead (temporal filtering / TPL windows), so parity is defined per sub-stream: the CUDA-backed
encode of rank r's sub-clip must be bit-identical to the CPU reference encode of the same sub-clip (checked with --verify).

    python tools/shard_encode.py --width 1920 --height 1080 --frames 256 --gop 32 --gpus 2 [--verify] [--sequential]

--sequential runs the ranks one after the other on GPU 0 (functional check on a 1-GPU box); otherwise rank r uses GPU r and
all ranks run concurrently, each bound to its own group of host cores.  Prints one JSON line.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "svt-av1_b200"), os.path.join(ROOT, "tools"), ROOT):
    sys.path.insert(0, p)
import encode_compare as ec  # noqa: E402
import make_yuv  # noqa: E402
import sharding  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=640); ap.add_argument("--height", type=int, default=360)
    ap.add_argument("--bits", type=int, default=8); ap.add_argument("--preset", type=int, default=8)
    ap.add_argument("--qp", type=int, default=43); ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--gop", type=int, default=32, help="frames per closed GOP (--keyint gop-1)")
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--sequential", action="store_true")
    ap.add_argument("--verify", action="store_true", help="also encode every sub-clip with the CPU reference and compare md5")
    ap.add_argument("--decode", action="store_true",
                    help="decode the spliced stream with the reference's own decoder (oracle/_ref/app/SvtAv1DecApp) and compare with "
                         "the ranks' reconstructions spliced the same way")
    ap.add_argument("--variant", default="cuda_simd")
    ap.add_argument("--workdir", default=None)
    a = ap.parse_args()
    wd = a.workdir or tempfile.mkdtemp(prefix="svtshard_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    os.makedirs(wd, exist_ok=True)
    clip = os.path.join(wd, "clip.yuv")
    make_yuv.write_clip(clip, a.width, a.height, a.frames, a.bits)
    fbytes = (a.width * a.height + 2 * (a.width // 2) * (a.height // 2)) * (2 if a.bits > 8 else 1)
    subs, owners = sharding.split_clip(clip, fbytes, a.gop, a.gpus, wd)
    import bench
    parts = [bench.cpu_partition(a.gpus, r) if not a.sequential else None for r in range(a.gpus)]

    def launch(r, variant, tag):
        base = variant.split(":")[0]
        env = dict(os.environ)
        for k in list(env):
            if k.startswith("SVT_CUDA"):
                del env[k]
        if base.startswith("cuda"):
            env["SVT_CUDA"] = "1"
            env["SVT_CUDA_DEVICE"] = "0" if a.sequential else str(r)
        n = os.path.getsize(subs[r]) // fbytes
        ivf = os.path.join(wd, "%s_r%d.ivf" % (tag, r))
        cmd = [ec.APPS[base], "-i", subs[r], "-w", str(a.width), "-h", str(a.height), "--fps", "30", "--preset", str(a.preset),
               "--rc", "0", "-q", str(a.qp), "-n", str(n), "--keyint", str(a.gop - 1), "-b", ivf] + \
              (["-o", ivf[:-4] + ".rec"] if a.decode and tag == "gpu" else []) + \
              (["--input-depth", str(a.bits)] if a.bits != 8 else []) + (["--lp", str(len(parts[r]))] if parts[r] else [])

        def pre():
            if parts[r]:
                os.sched_setaffinity(0, parts[r])
        return subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, preexec_fn=pre), ivf

    def run_all(variant, tag):
        t0 = time.time()
        ivfs = []
        if a.sequential:
            for r in range(a.gpus):
                p, ivf = launch(r, variant, tag)
                out = p.communicate()[0]
                assert p.returncode == 0, out.decode(errors="replace")[-2000:]
                ivfs.append(ivf)
        else:
            ps = [launch(r, variant, tag) for r in range(a.gpus)]
            for p, ivf in ps:
                out = p.communicate()[0]
                assert p.returncode == 0, out.decode(errors="replace")[-2000:]
                ivfs.append(ivf)
        return time.time() - t0, ivfs

    secs, ivfs = run_all(a.variant, "gpu")
    hdr, _ = sharding.ivf_packets(ivfs[0])
    packets = [sharding.ivf_packets(f)[1] for f in ivfs]
    whole = sharding.splice_gops(packets, owners, a.gop, a.frames)
    out = os.path.join(wd, "spliced.ivf")
    sharding.write_ivf(out, hdr, whole)
    res = {"frames": a.frames, "gop": a.gop, "ranks": a.gpus, "gops_per_rank": [len(o) for o in owners], "wall_s": round(secs, 3),
           "fps_wall_incl_init": round(a.frames / secs, 3), "packets": len(whole), "spliced_ivf": out,
           "spliced_md5": hashlib.md5(open(out, "rb").read()).hexdigest(), "rank_md5": [ec.md5(f) for f in ivfs]}
    if a.decode:
        dec = os.path.join(wd, "spliced_dec.yuv")
        app = os.path.join(ROOT, "oracle", "_ref", "app", "SvtAv1DecApp")
        p = subprocess.run([app, "-i", out, "-o", dec] + (["-bit-depth", str(a.bits)] if a.bits != 8 else []),
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        assert p.returncode == 0 and os.path.exists(dec), p.stdout.decode(errors="replace")[-1500:]
        # the ranks' reconstructions (display order inside every closed GOP), spliced GOP by GOP like the packets
        recs = [open(f[:-4] + ".rec", "rb").read() for f in ivfs]
        frames = [[r[i:i + fbytes] for i in range(0, len(r), fbytes)] for r in recs]
        want = b"".join(sharding.splice_gops(frames, owners, a.gop, a.frames))
        res["decoded_md5"] = ec.md5(dec)
        res["spliced_recon_md5"] = hashlib.md5(want).hexdigest()
        res["decode_matches_recon"] = res["decoded_md5"] == res["spliced_recon_md5"]
    if a.verify:
        _, refs = run_all("ref_simd" if a.variant.endswith("simd") else "ref_c", "cpu")
        res["cpu_rank_md5"] = [ec.md5(f) for f in refs]
        res["parity"] = res["cpu_rank_md5"] == res["rank_md5"]
    print(json.dumps(res))
    return 0 if (not a.verify or res["parity"]) and (not a.decode or res["decode_matches_recon"]) else 1


if __name__ == "__main__":
    sys.exit(main())
