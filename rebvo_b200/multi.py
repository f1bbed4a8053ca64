"""Multi-GPU plumbing for the replay bench: one process per GPU, independent sequences per rank (the tracker is a
recurrence over frames, SURVEY.md 8(e): replicas only), torch.distributed for the barrier and the max-over-ranks
time.  Backend nccl on GPUs, gloo in the CPU tests."""
import os


def rank_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def stream_seed(rank, base=7):
    """Each rank replays its own sequence (seed) -- frames are never shared between ranks."""
    return base + rank


def init(backend, device=None):
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29517")
    if device is not None:
        dist.init_process_group(backend, device_id=device)
    else:
        dist.init_process_group(backend)
    return dist


def max_over_ranks(dist, values, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks."""
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def aggregate_fps(frames_per_rank, world, t_max_ms):
    """Whole-job frames/s: every rank processed frames_per_rank frames, the job took the slowest rank's time."""
    return frames_per_rank * world / (t_max_ms * 1e-3)
