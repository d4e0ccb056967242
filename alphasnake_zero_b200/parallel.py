"""Multi-GPU plumbing: self-play shards by game, one process per GPU, no collective on the hot path.
The only communication is per generation: the weight broadcast from the training rank (alpha_snake_zero_trainer.py:52-57
hands the same nnet to every self-play game), the hand-off of the sampled training batch and the six log counters.
Every collective takes tensors on the rank's CUDA device under the nccl backend (NVLink / NVSwitch) and CPU tensors under
gloo (the CPU tests)."""
import random

import numpy as np
import torch
import torch.distributed as dist

from .utils.alpha_nnet import flatten_weights


def shard_range(total_games, rank, world):
    """rank r owns games [r*G/R, (r+1)*G/R) (contiguous, sizes differ by at most one)."""
    lo = total_games * rank // world
    hi = total_games * (rank + 1) // world
    return lo, hi


def collective_device(device=None):
    """device the collectives' tensors must live on: the rank's CUDA device for nccl, the CPU for gloo"""
    if device is not None:
        return torch.device(device)
    if dist.is_initialized() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def weights_checksum(weights):
    """order-sensitive 64-bit checksum of every weight / BN buffer (bit patterns, not values)"""
    acc = np.uint64(0)
    with np.errstate(over="ignore"):
        for i, a in enumerate(flatten_weights(weights)):
            u = np.ascontiguousarray(a, dtype=np.float32).reshape(-1).view(np.uint32).astype(np.uint64)
            k = np.arange(1, u.size + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(i)
            acc = acc + (u * k).sum(dtype=np.uint64)
    return int(acc)


def broadcast_weights(target, src=0, device=None):
    """One flat fp32 message (~5 MB at 11x11) from rank `src` to every rank.

    target: an AlphaNNet (its weight dictionary is overwritten and the device copies behind `v` -- the native network's
    operands or the PyTorch parameters -- are refreshed through AlphaNNet.set_weights / asz_net_update_weights), or a bare
    weight dictionary (updated in place).  Returns the weight dictionary."""
    weights = target.weights if hasattr(target, "weights") else target
    arrs = flatten_weights(weights)
    flat = torch.from_numpy(np.concatenate([np.asarray(a, np.float32).reshape(-1) for a in arrs])).to(collective_device(device))
    dist.broadcast(flat, src=src)
    if dist.get_rank() != src:
        host = flat.cpu().numpy()
        o = 0
        for a in arrs:
            a[...] = host[o:o + a.size].reshape(a.shape)
            o += a.size
        if hasattr(target, "set_weights"):
            target.set_weights(weights)
    return weights


def weights_equal_all_ranks(weights, device=None):
    """True when every rank holds bit-identical weights (min and max of the checksum agree)"""
    c = weights_checksum(weights)
    # two 31-bit halves + the rest: float64 all_reduce is exact on values below 2^53, int64 MIN/MAX is not available everywhere
    parts = [c & 0x7fffffff, (c >> 31) & 0x7fffffff, c >> 62]
    t = torch.tensor(parts + [-p for p in parts], dtype=torch.float64, device=collective_device(device))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.cpu().tolist()
    return all(t[i] == -t[i + 3] for i in range(3))


def reduce_counters(local_sums, games_local, dst=0, device=None):
    """Per-game averages of the six log counters over all ranks (mp_game_runner.py:71-76 on the union of shards)."""
    t = torch.tensor(list(local_sums) + [float(games_local)], dtype=torch.float64, device=collective_device(device))
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    if dist.get_rank() != dst:
        return None
    t = t.cpu()
    return (t[:-1] / t[-1]).tolist()


def gather_sampled_batch(local_count, local_gather, plane_shape, batch_size=2048, max_batches=5, dst=0, rng=random, device=None):
    """alpha_snake_zero_trainer.py:62-77 over the union of every rank's records, moving only the SAMPLED records.

    local_count: records this rank holds; local_gather(idx) -> (X [k, *plane_shape], V [k, 3]) float32 tensors for local
    indices (Engine.records_gather(idx, mirror=False): one gather kernel).  Rank dst draws `samples` distinct GLOBAL indices
    like the reference's random.sample, broadcasts them, every rank gathers the ones it owns and they meet on dst in one
    padded tensor gather.  Returns (X, V, batch_size) on dst (NOT mirrored: the caller mirrors once, on the device) and
    (None, None, batch_size) elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = collective_device(device)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[rank] = local_count
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    counts = counts.cpu().tolist()
    total = int(sum(counts))
    batches = min(max_batches, total // batch_size)
    samples = batch_size * batches
    if samples > total or samples == 0:             # :67-69
        batch_size = total
        samples = total
    idx = torch.zeros(max(samples, 1), dtype=torch.int64, device=dev)
    if rank == dst and samples:
        idx[:samples] = torch.tensor(rng.sample(range(total), samples), dtype=torch.int64)
    dist.broadcast(idx, src=dst)
    idx = idx[:samples].cpu()
    lo = int(sum(counts[:rank]))
    mine = idx[(idx >= lo) & (idx < lo + counts[rank])] - lo
    k = int(mine.numel())
    plane = int(np.prod(plane_shape))
    # every rank's share, padded to the largest share (tensor gather needs equal shapes)
    ks = torch.zeros(world, dtype=torch.int64, device=dev)
    ks[rank] = k
    dist.all_reduce(ks, op=dist.ReduceOp.SUM)
    ks = ks.cpu().tolist()
    kmax = max(max(ks), 1)
    buf = torch.zeros(kmax, plane + 3, dtype=torch.float32, device=dev)
    if k:
        X, V = local_gather(mine)
        buf[:k, :plane] = X.reshape(k, plane).to(dev)
        buf[:k, plane:] = V.to(dev)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, out, dst=dst)
    if rank != dst:
        return None, None, batch_size
    allb = torch.cat([out[r][:ks[r]] for r in range(world)])
    return allb[:, :plane].reshape(-1, *plane_shape).contiguous(), allb[:, plane:].contiguous(), batch_size
