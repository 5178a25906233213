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
