"""CPU, world size 2, gloo: the multi-GPU plumbing (game sharding, weight broadcast, record gather)."""
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
    w = init_weights((21, 21, 3), seed=100 + rank)          # every rank starts with different weights
    parallel.broadcast_weights(w, src=0)
    ref = init_weights((21, 21, 3), seed=100)
    same = all(np.array_equal(a, b) for a, b in zip(flatten_weights(w), flatten_weights(ref)))
    lo, hi = parallel.shard_range(1001, rank, world)
    recs = [np.full((21, 21, 3), rank, np.float32)] * (3 + rank)
    vals = [np.full(3, rank, np.float32)] * (3 + rank)
    R, V = parallel.gather_records(recs, vals, dst=0)
    avg = parallel.reduce_counters([1.0 * (hi - lo), 2.0 * (hi - lo), 0, 0, 0, 10.0 * (hi - lo)], hi - lo, dst=0)
    q.put((rank, same, lo, hi, len(R), len(V), avg))
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
    assert res[0][4] == res[0][5] == 3 + 4 and res[1][4] == 0
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
