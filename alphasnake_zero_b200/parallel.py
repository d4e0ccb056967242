"""Multi-GPU plumbing: self-play shards by game, one process per GPU, no collective on the hot path.
The only communication is per generation: the weight broadcast from the training rank and the gather of the
training records and log counters.  Works with the nccl backend on GPUs and with gloo on CPU (tests)."""
import numpy as np
import torch
import torch.distributed as dist

from .utils.alpha_nnet import flatten_weights


def shard_range(total_games, rank, world):
    """rank r owns games [r*G/R, (r+1)*G/R) (contiguous, sizes differ by at most one)."""
    lo = total_games * rank // world
    hi = total_games * (rank + 1) // world
    return lo, hi


def _weight_arrays(weights):
    return flatten_weights(weights)


def broadcast_weights(weights, src=0, device=None):
    """In-place broadcast of every weight / BN buffer of an AlphaNNet weight dict (one flat fp32 message, ~5 MB)."""
    arrs = _weight_arrays(weights)
    flat = torch.from_numpy(np.concatenate([a.reshape(-1).astype(np.float32) for a in arrs]))
    if device is not None:
        flat = flat.to(device)
    dist.broadcast(flat, src=src)
    flat = flat.cpu().numpy()
    o = 0
    for a in arrs:
        a[...] = flat[o:o + a.size].reshape(a.shape)
        o += a.size
    return weights


def gather_records(records, values, dst=0):
    """Training pairs of every rank on rank dst (lists of numpy arrays); other ranks get ([], [])."""
    world, rank = dist.get_world_size(), dist.get_rank()
    payload = (np.array(records, np.float32), np.array(values, np.float32))
    out = [None] * world if rank == dst else None
    dist.gather_object(payload, out, dst=dst)
    if rank != dst:
        return [], []
    recs, vals = [], []
    for r, v in out:
        recs += list(r)
        vals += list(v)
    return recs, vals


def reduce_counters(local_sums, games_local, dst=0):
    """Per-game averages of the six log counters over all ranks (mp_game_runner.py:71-76 on the union of shards)."""
    t = torch.tensor(list(local_sums) + [float(games_local)], dtype=torch.float64)
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    if dist.get_rank() != dst:
        return None
    return (t[:-1] / t[-1]).tolist()
