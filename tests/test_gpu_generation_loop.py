"""One generation of the reference's training loop (alpha_snake_zero_trainer.py:52-91) and one pit (pit.py:28-35) on the
drop-in classes, with tiny settings: self-play with the search agent -> log counters -> sampled records, mirrored
(:93-100) -> nnet.train -> save / reload -> the new generation plays the old one."""
import os
from random import sample, seed as rseed

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_one_generation_and_a_pit(tmp_path, monkeypatch):
    import torch
    from alphasnake_zero_b200.utils.agent import Agent
    from alphasnake_zero_b200.utils.alpha_nnet import AlphaNNet
    from alphasnake_zero_b200.utils.mp_game_runner import MPGameRunner
    from alphasnake_zero_b200.utils import pit_agent, pit_mp_game_runner
    monkeypatch.chdir(tmp_path)
    rseed(0); torch.manual_seed(0)
    nnet = AlphaNNet(input_shape=(21, 21, 3), seed=3)                     # train.py:32
    nnet = nnet.copy_and_compile()                                        # alpha_snake_zero_trainer.py:34
    Alice = Agent(nnet, 2, True, 4, 16)                                   # :52 (depth 4, breadth 16 instead of 8, 128)
    gr = MPGameRunner(11, 11, 4, 9, 48, verbose=False)                    # :53
    rewards = gr.run(Alice)                                               # :54, to completion
    assert len(rewards) == 48 and all(r is not None and len(r) == 4 for r in rewards)
    for r in rewards:                                                     # at most one winner, everybody else lost
        assert sorted(x for x in r if x is not None)[:-1].count(1.0) == 0 and r.count(1.0) <= 1 and set(r) <= {1.0, -1.0}
    log_list = [gr.wall_collision, gr.body_collision, gr.head_collision, gr.starvation, gr.food_eaten, gr.game_length]
    assert all(np.isfinite(x) and x >= 0 for x in log_list) and gr.game_length > 1
    deaths = gr.wall_collision + gr.body_collision + gr.head_collision + gr.starvation
    assert 3.0 <= deaths <= 4.0                                           # 3 or 4 snakes die per game
    n = len(Alice.records)
    assert n == len(Alice.values) and n > 48 * 4
    batch_size = min(256, n)
    idx = sample(range(n), batch_size)                                    # :71
    X = [Alice.records[i] for i in idx]
    V = [Alice.values[i] for i in idx]
    assert X[0].shape == (21, 21, 3) and X[0].dtype == np.float32 and V[0].shape == (3,)
    Alice.clear()                                                         # :75
    assert len(Alice.records) == 0
    X += list(np.flip(X, axis=2))                                         # mirror_states, :93-97
    V += list(np.flip(V, axis=1))                                         # mirror_values, :99-100
    before = nnet.v(X[:32])
    new = nnet.copy_and_compile(learning_rate=1e-3)                       # :79
    new.train(X, V, epochs=2, batch_size=batch_size)                      # :81
    new = new.copy_and_compile()                                          # :83
    after = new.v(X[:32])
    assert np.isfinite(after).all() and np.abs(after - before).max() > 1e-4        # the weights moved
    new.save("Test1")                                                     # :91 -> models/Test1.h5 (Keras layout, utils/h5lite.py)
    assert os.path.exists("models/Test1.h5")
    loaded = AlphaNNet(model_name="models/Test1.h5")                      # train.py:35 passes "models/<name><generation>.h5"
    np.testing.assert_allclose(loaded.v(X[:32]), after, rtol=0, atol=1e-6)
    np.testing.assert_allclose(AlphaNNet(model_name="models/Test1").v(X[:32]), after, rtol=0, atol=1e-6)
    # pit.py:28-35: the new generation (snake ids 0..1) against the old one, 2 vs 2
    A, B = pit_agent.Agent(loaded), pit_agent.Agent(nnet)
    winners = pit_mp_game_runner.MPGameRunner(11, 11, 4, 9, 24).run(A, B, 2)
    assert len(winners) == 24 and all(w is None or 0 <= w < 4 for w in winners)
    assert sum(w is not None for w in winners) > 0


def test_device_records_equal_host_lists_and_mirrored_batch():
    """Agent.records / Agent.values live in HBM (asz_records_*): after every root turn the store must hold exactly what the
    reference's host lists would hold (agent.py:93-97: the root state and root Q row of every live snake, in ids order), and
    the sampled + mirrored training batch (alpha_snake_zero_trainer.py:70-77, 93-100) must equal numpy.flip of the host path."""
    import torch
    from alphasnake_zero_b200.utils.agent import Agent, StubNet
    from alphasnake_zero_b200.utils.mp_game_runner import MPGameRunner
    G, S = 96, 4
    inner = Agent(StubNet(), 2, True, 4, 16, records_capacity=64)         # 64 rows: the store has to grow several times
    host_records, host_values = [], []

    class Shadow:
        """forwards to the real agent, then rebuilds the reference's host lists from the same engine state"""
        def __getattr__(self, k):
            return getattr(inner, k)

        def make_moves(self, games, ids):
            moves = inner.make_moves(games, ids)
            eng = games.engine
            q = eng._wrap(eng.L.asz_search_root_q(eng.h), (eng.G, 8, 3), torch.float32).cpu().numpy()
            planes, rows = eng.encode_rows()
            ph = planes.cpu().numpy()
            row_of = {(int(r) // 8, int(r) % 8): i for i, r in enumerate(rows)}
            for g, s in ids:
                host_records.append(ph[row_of[(g, s)]].copy())
                host_values.append(q[g, s].copy())
            return moves

    gr = MPGameRunner(11, 11, S, 3, G, verbose=False, seed=11)
    gr.run(Shadow(), max_turns=12)
    eng = gr.engine
    n = len(inner.records)
    assert n == len(host_records) == len(inner.values) > G * 4
    pl, va, ids, turns = eng.records_views()
    order = np.lexsort((ids.cpu().numpy(), turns.cpu().numpy()))          # per turn, ids order = ascending game*8 + snake
    assert np.array_equal(pl.cpu().numpy()[order].view(np.uint32), np.array(host_records).view(np.uint32))
    assert np.array_equal(va.cpu().numpy()[order].view(np.uint32), np.array(host_values).view(np.uint32))
    assert np.array_equal(inner.records[5], pl[5].cpu().numpy()) and inner.values[-1].shape == (3,)
    # the trainer's batch: sample without replacement, then mirrored copies appended (states flipped along the width axis,
    # values reversed)
    rng = np.random.default_rng(0)
    idx = rng.choice(n, size=min(n, 700), replace=False)
    X, V = eng.records_gather(idx, mirror=True)
    Xh, Vh = pl.cpu().numpy()[idx], va.cpu().numpy()[idx]
    want_X = np.concatenate([Xh, np.flip(Xh, axis=2)])                    # mirror_states, :93-97
    want_V = np.concatenate([Vh, np.flip(Vh, axis=1)])                    # mirror_values, :99-100
    assert np.array_equal(X.cpu().numpy().view(np.uint32), want_X.view(np.uint32))
    assert np.array_equal(V.cpu().numpy().view(np.uint32), want_V.view(np.uint32))
    X1, V1 = eng.records_gather(idx[:10], mirror=False)
    assert X1.shape == (10, 21, 21, 3) and np.array_equal(X1.cpu().numpy(), Xh[:10])
    Xb, Vb, bs = inner.sample_training_batch(batch_size=256, max_batches=5)
    assert bs == 256 and Xb.shape[0] == 2 * 256 * min(5, n // 256) and Vb.shape == (Xb.shape[0], 3) and Xb.is_cuda
    assert torch.equal(Xb[Xb.shape[0] // 2:], torch.flip(Xb[:Xb.shape[0] // 2], dims=[2]))
    # the device batch feeds AlphaNNet.train without a host round trip
    from alphasnake_zero_b200.utils.alpha_nnet import AlphaNNet
    net = AlphaNNet(input_shape=(21, 21, 3), seed=3).copy_and_compile(learning_rate=1e-3)
    before = net.v(Xb[:16])
    net.train(Xb, Vb, epochs=1, batch_size=bs)
    assert np.abs(net.v(Xb[:16]) - before).max() > 1e-5
    inner.clear()
    assert len(inner.records) == 0 and eng.records_count() == 0
