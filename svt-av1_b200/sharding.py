"""Multi-GPU host logic of the hot path: independent streams / mini-GOPs are sharded one group per
rank with no data-path exchange; the only collectives are the barrier around the timed region and the MAX-reduce of
the per-rank device time.  Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def assign_streams(n_streams, world):
    """Stream indices owned by each rank: contiguous, balanced, covering every stream exactly once."""
    base, extra = divmod(n_streams, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append(list(range(start, start + n)))
        start += n
    return out


def stream_seed(base_seed, rank):
    """Every rank encodes a different synthetic stream (config 4: one stream per GPU)."""
    return base_seed + rank


def reduce_max_ms(ms, device="cpu"):
    """Whole-job step time = the slowest rank's device time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(ms)
    t = torch.tensor([float(ms)], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_fps(frames_per_rank, ms_max, world):
    """value = units all ranks processed / max-over-ranks time."""
    return frames_per_rank * world / (ms_max / 1e3)


# ---- one stream over several GPUs (BASELINE configs[4]): closed GOPs round-robin over the ranks ------------
def assign_gops(n_gops, world):
    """GOP indices of each rank: round-robin (rank r owns r, r + world, ...), so that every rank's sub-stream spans the whole
    clip and the ranks finish together; every GOP is owned exactly once."""
    return [list(range(r, n_gops, world)) for r in range(world)]


def split_clip(clip_path, frame_bytes, gop_frames, world, out_dir, tag="shard"):
    """Writes rank r's sub-clip (its GOPs, concatenated in order) next to the others; returns (paths, gops per rank).
    The last GOP may be shorter."""
    import os
    total = os.path.getsize(clip_path) // frame_bytes
    n_gops = -(-total // gop_frames)
    owners = assign_gops(n_gops, world)
    paths = [os.path.join(out_dir, "%s_r%d_of%d.yuv" % (tag, r, world)) for r in range(world)]
    outs = [open(p, "wb") for p in paths]
    with open(clip_path, "rb") as f:
        for g in range(n_gops):
            data = f.read(frame_bytes * gop_frames)
            outs[g % world].write(data)
    for o in outs:
        o.close()
    return paths, owners


def ivf_packets(path):
    """(header bytes, [packet payloads]) of an IVF file."""
    import struct
    d = open(path, "rb").read()
    hl = struct.unpack("<H", d[6:8])[0]
    pos, out = hl, []
    while pos < len(d):
        n = struct.unpack("<I", d[pos:pos + 4])[0]
        out.append(d[pos + 12:pos + 12 + n])
        pos += 12 + n
    return d[:hl], out


def splice_gops(rank_packets, owners, gop_frames, total_frames):
    """Interleaves the per-rank packet lists (each rank's packets are its GOPs back to back, gop_frames packets per GOP in
    decode order, every GOP opening with a key frame) into stream order.  Returns the packet list of the whole stream."""
    n_gops = -(-total_frames // gop_frames)
    cursor = [0] * len(rank_packets)
    out = []
    for g in range(n_gops):
        r = g % len(rank_packets)
        n = min(gop_frames, total_frames - g * gop_frames)
        out += rank_packets[r][cursor[r]:cursor[r] + n]
        cursor[r] += n
    assert all(c == len(p) for c, p in zip(cursor, rank_packets)), "packet count of a rank does not match its GOPs"
    return out


def write_ivf(path, header, packets):
    import struct
    with open(path, "wb") as f:
        h = bytearray(header)
        h[24:28] = struct.pack("<I", len(packets))
        f.write(bytes(h))
        for i, p in enumerate(packets):
            f.write(struct.pack("<IQ", len(p), i))
            f.write(p)
