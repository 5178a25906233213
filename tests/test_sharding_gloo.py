"""N>1 host logic on CPU: world_size-2 gloo processes (rendezvous on 127.0.0.1)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "svt-av1_b200"))


def _worker(rank, world, port, q):
    import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.assign_streams(8, world)[rank]
    ms = 10.0 + 5.0 * rank  # rank 1 is slower
    dist.barrier()
    mx = sharding.reduce_max_ms(ms)
    # every stream is owned exactly once
    owned = [torch.zeros(8, dtype=torch.int64) for _ in range(world)]
    mask = torch.zeros(8, dtype=torch.int64)
    mask[mine] = 1
    dist.all_gather(owned, mask)
    total = torch.stack(owned).sum(0)
    q.put((rank, mine, mx, total.tolist(), sharding.stream_seed(1234, rank)))
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_reduce():
    import sharding
    assert sharding.assign_streams(8, 1) == [list(range(8))]
    assert sharding.assign_streams(7, 2) == [[0, 1, 2, 3], [4, 5, 6]]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 1, 2, 3] and res[1][1] == [4, 5, 6, 7]
    assert res[0][2] == res[1][2] == 15.0  # max over ranks
    assert res[0][3] == [1] * 8
    assert res[0][4] != res[1][4]
    assert sharding.aggregate_fps(80, 15.0, 2) == 160 / 0.015


def test_gop_sharding_round_robin_and_splice(tmp_path):
    """One stream over N encoder instances: every GOP owned once, round-robin; splitting a clip and splicing
    the per-rank packet lists back restores stream order (ragged last GOP included)."""
    import sharding
    assert sharding.assign_gops(7, 3) == [[0, 3, 6], [1, 4], [2, 5]]
    for world in (1, 2, 3, 8):
        owners = sharding.assign_gops(11, world)
        assert sorted(g for o in owners for g in o) == list(range(11))
    fb, gop, frames = 16, 4, 18  # 18 "frames" of 16 bytes, GOPs of 4 (the last one has 2)
    clip = tmp_path / "clip.bin"
    clip.write_bytes(b"".join(bytes([i]) * fb for i in range(frames)))
    for world in (1, 2, 3):
        paths, owners = sharding.split_clip(str(clip), fb, gop, world, str(tmp_path), tag="t%d" % world)
        per_rank = []
        for p in paths:
            d = open(p, "rb").read()
            per_rank.append([d[i:i + fb] for i in range(0, len(d), fb)])  # one "packet" per frame
        whole = sharding.splice_gops(per_rank, owners, gop, frames)
        assert whole == [bytes([i]) * fb for i in range(frames)]
