"""Multi-GPU plumbing of the encode path: independent closed-GOP chunks, one chunk stream
per GPU / rank, NO collective on the data path.  torch.distributed is used only for the barrier around
the timed region and the max-over-ranks reduction of the elapsed time (bench.py) -- gloo on CPU in the
tests, nccl on the GPU box."""
import os


def plan_chunks(n_frames, keyint):
    """av1an-style chunking: closed GOPs of at most `keyint` frames -> list of (first_frame, n_frames)."""
    if n_frames <= 0 or keyint <= 0:
        return []
    return [(s, min(keyint, n_frames - s)) for s in range(0, n_frames, keyint)]


def chunks_of_rank(chunks, rank, world):
    """Chunk c is encoded by rank c mod world (the same static round-robin the av1an front end uses)."""
    return [c for i, c in enumerate(chunks) if i % world == rank]


def host_threads_per_rank(world_on_node=None):
    """Entropy-coding threads per rank: the node's cores divided among the ranks that share it."""
    cores = os.cpu_count() or 1
    w = world_on_node or int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    return max(1, cores // max(1, w))


def max_over_ranks(value, dist=None, device="cpu"):
    """Elapsed-time reduction of the bench contract: every rank passes its own time, all get the maximum."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def concat_in_order(per_rank_results, n_chunks, world):
    """per_rank_results[r] = list of chunk payloads of rank r in its own order -> payloads in chunk order."""
    out = []
    idx = [0] * world
    for c in range(n_chunks):
        r = c % world
        out.append(per_rank_results[r][idx[r]])
        idx[r] += 1
    return out
