"""CPU, world size 2, gloo: the multi-GPU plumbing (game sharding, weight broadcast + equality check, hand-off of the sampled
training batch, log counters)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from alphasnake_zero_b200 import parallel
    from alphasnake_zero_b200.utils.alpha_nnet import init_weights, flatten_weights
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import random
    import torch
    w = init_weights((21, 21, 3), seed=100 + rank)          # every rank starts with different weights
    differ_before = not parallel.weights_equal_all_ranks(w)
    parallel.broadcast_weights(w, src=0)
    ref = init_weights((21, 21, 3), seed=100)
    same = all(np.array_equal(a, b) for a, b in zip(flatten_weights(w), flatten_weights(ref)))
    same = same and parallel.weights_equal_all_ranks(w) and differ_before
    lo, hi = parallel.shard_range(1001, rank, world)
    # records of this rank: record i of rank r is filled with the value 1000 * r + i (so the origin of every sampled row shows)
    n_local = 300 + 50 * rank
    planes = torch.arange(n_local, dtype=torch.float32).view(-1, 1, 1, 1).expand(n_local, 21, 21, 3) + 1000.0 * rank
    values = torch.stack([planes[:, 0, 0, 0], planes[:, 0, 0, 0] + 0.25, planes[:, 0, 0, 0] + 0.5], 1)
    X, V, bs = parallel.gather_sampled_batch(n_local, lambda idx: (planes[idx], values[idx]), (21, 21, 3), batch_size=128,
                                             max_batches=5, dst=0, rng=random.Random(5))
    got = None
    if rank == 0:
        tags = X[:, 0, 0, 0]
        ok = bool((X == tags.view(-1, 1, 1, 1)).all()) and bool((V[:, 0] == tags).all()) and bool((V[:, 2] == tags + 0.5).all())
        got = (tuple(X.shape), bs, len(set(tags.tolist())), int((tags >= 1000).sum()), ok)
    avg = parallel.reduce_counters([1.0 * (hi - lo), 2.0 * (hi - lo), 0, 0, 0, 10.0 * (hi - lo)], hi - lo, dst=0)
    q.put((rank, same, lo, hi, got, X is None, avg))
    dist.destroy_process_group()


def test_sharding_broadcast_gather_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), "weights differ after the broadcast"
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 500, 500, 1001)     # contiguous, disjoint, complete
    # alpha_snake_zero_trainer.py:62-70 over the union of both ranks' records (300 + 350): 5 batches of 128 distinct records,
    # some from each rank, every row intact; nothing arrives on the other rank
    shape, bs, distinct, from_rank1, intact = res[0][4]
    assert shape == (640, 21, 21, 3) and bs == 128 and distinct == 640 and 200 < from_rank1 < 450 and intact
    assert res[0][5] is False and res[1][5] is True
    assert np.allclose(res[0][6], [1.0, 2.0, 0, 0, 0, 10.0])


def test_shard_range_properties():
    from alphasnake_zero_b200.parallel import shard_range
    for G in (1, 7, 8, 32768, 65537):
        for R in (1, 2, 4, 8):
            spans = [shard_range(G, r, R) for r in range(R)]
            assert spans[0][0] == 0 and spans[-1][1] == G
            assert all(spans[i][1] == spans[i + 1][0] for i in range(R - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
