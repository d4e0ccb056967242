"""GPU parity of the search kernels (through the C ABI):
 (1) replaying the in-tree moves the unmodified reference sampled: visit counts, key sets and ages bit-exact,
     Q/W within float32 reassociation (the reference backs up sequentially, the GPU with atomics);
 (2) native GPU sampling recorded as a trace, replayed by the CPU oracle."""
import numpy as np
import pytest

from tests.helpers import assert_dump_equal, load
from tests.test_gpu_env import _engine, init_dump

pytestmark = pytest.mark.gpu

MCTS = ["11x11x4_train", "11x11x4_eval", "7x7x4_dec9", "19x19x8", "11x11x4_b5"]


def rows_of(eng_or_dumps, G, S, alive_fn):
    rows = []
    for g in range(G):
        for s in range(S):
            if alive_fn(g, s):
                rows.append((g, s))
    return rows


@pytest.mark.parametrize("name", MCTS)
def test_search_replay_against_reference(name):
    import torch
    z = load("mcts_%s.npz" % name)
    side, S, dec, G = int(z["H"]), int(z["S"]), int(z["health_dec"]), int(z["G"])
    D, breadth, training = int(z["D"]), int(z["breadth"]), bool(z["training"])
    eng = _engine(side=side, snakes=S, health_dec=dec, games=G, seed=5, max_depth=D, max_breadth=breadth,
                  softmax_base=float(z["base"]), training=training, table_log2=16)
    for gi in range(G):
        nf = int(z["init_nfood"][gi])
        eng.set_state(gi, init_dump(side, S, z["init_start"][gi], z["init_last"][gi], z["init_food"][gi][:nf]))
    info = eng.search_info()
    assert info["P"] == min(8, breadth) and info["epochs"] == breadth // min(8, breadth)
    for t in range(int(z["n_turns"])):
        live = z["t%d_live" % t].tolist()
        states = {}
        for j, gid in enumerate(live):
            want = {k: z["t%d_before_%s" % (t, k)][j] for k in ("snake", "owner", "dist", "food", "counters")}
            got = eng.get_state(gid)
            assert_dump_equal(got, want, "root game %d turn %d" % (gid, t))
            assert got["counters"][7] == 0
            states[gid] = want
        for gid in range(G):
            if gid not in live:
                assert eng.get_state(gid)["counters"][7] == 1   # finished games stay finished
        ids = z["t%d_ids" % t]
        want_moves = z["t%d_root_moves" % t]
        root_trace = np.full((G, 8), 255, np.uint8)
        for (gid, sid), m in zip(ids, want_moves):
            root_trace[gid, sid] = m
        tree = torch.from_numpy(np.ascontiguousarray(z["t%d_tree" % t])).cuda()
        q, mv = eng.search(value_fn=None, trace=tree, trace_mode=1, root_trace=torch.from_numpy(root_trace).cuda())
        q = q.cpu().numpy(); mv = mv.cpu().numpy()
        got_q = np.array([q[gid, sid] for gid, sid in ids])
        np.testing.assert_allclose(got_q, z["t%d_root_q" % t], rtol=0, atol=3e-6)
        got_mv = np.array([mv[gid, sid] for gid, sid in ids])
        if training:
            assert np.array_equal(got_mv, want_moves)
        else:
            srt = np.sort(z["t%d_root_q" % t], axis=1)
            safe = (srt[:, 2] - srt[:, 1]) > 1e-5
            assert np.array_equal(got_mv[safe], want_moves[safe])
        assert (mv != 255).sum() == len(ids)
        tab = eng.table()
        assert np.array_equal(tab["keys"], z["t%d_tab_keys" % t]), "key sets differ at turn %d" % t
        assert np.array_equal(tab["N"], z["t%d_tab_N" % t]), "visit counts differ at turn %d" % t
        assert np.array_equal(tab["age"], z["t%d_tab_age" % t])
        np.testing.assert_allclose(tab["W"], z["t%d_tab_W" % t], rtol=0, atol=3e-5)
        np.testing.assert_allclose(tab["Q"], z["t%d_tab_Q" % t], rtol=0, atol=3e-6)
        st = eng.search_stats()
        assert st["evals"] == int(z["t%d_n_evals" % t])
        assert st["collisions"] == 0 and st["overflow"] == 0
        # root tic with the reference's moves and spawn cells
        actions = np.ones((G, 8), np.uint8)
        for (gid, sid), m in zip(ids, want_moves):
            actions[gid, sid] = m
        spawn = np.full(G, -1, np.int32)
        for j, gid in enumerate(live):
            spawn[gid] = int(z["t%d_spawn" % t][j])
        eng.step(actions=torch.from_numpy(actions).cuda(), spawn_cells=torch.from_numpy(spawn).cuda(), spawn_mode=1,
                 tic=True, encode=False)
    eng.close()


@pytest.mark.parametrize("side,S,G,D,breadth,base,training,turns", [
    (11, 4, 64, 8, 32, 2.0, True, 6), (7, 4, 32, 4, 16, 100.0, False, 8), (19, 8, 8, 8, 16, 3.0, True, 3),
    (11, 2, 40, 6, 8, 10.0, True, 12)])
def test_native_search_against_oracle(side, S, G, D, breadth, base, training, turns):
    """The GPU samples with its own RNG and records the trace; the oracle replays it."""
    import torch
    from oracle import oracle as orc
    seed = 77
    eng = _engine(side=side, snakes=S, health_dec=1, games=G, seed=seed, max_depth=D, max_breadth=breadth,
                  softmax_base=base, training=training, table_log2=20)
    eng.reset()
    info = eng.search_info()
    games = []
    for gi in range(G):
        g = orc.OracleGame(side, side, S, 1); g.init_native(seed, gi, 0); g.set_ids(gi, 0); games.append(g)
    agent = orc.OracleAgent(base=base, training=training, max_depth=D, max_breadth=breadth)
    done = [False] * G
    for t in range(turns):
        tree = torch.full((info["epochs"], info["max_steps"], G * info["P"], S), 255, dtype=torch.uint8, device="cuda")
        q, mv = eng.search(value_fn=None, trace=tree, trace_mode=2)
        q = q.cpu().numpy(); mv = mv.cpu().numpy()
        live = [g for gi, g in enumerate(games) if not done[gi]]
        if not live:
            break
        ids = [(g_i, s) for g_i in range(G) if not done[g_i] for s in games[g_i].live_ids()]
        root_moves = np.array([mv[g_i, s] for g_i, s in ids], np.uint8)
        assert np.all(root_moves < 3)
        omv, oq = agent.make_moves(live, G, root_turn=t, tree_moves=np.ascontiguousarray(tree.cpu().numpy()),
                                   root_moves=root_moves.copy(), replay=True)
        got_q = np.array([q[g_i, s] for g_i, s in ids])
        np.testing.assert_allclose(got_q, oq, rtol=0, atol=5e-6)
        if training:
            assert np.array_equal(omv, root_moves)
        tab, otab = eng.table(), agent.table()
        oo = np.lexsort((otab["keys"][:, 1], otab["keys"][:, 0]))
        assert np.array_equal(tab["keys"], otab["keys"][oo]), "key sets differ at turn %d" % t
        assert np.array_equal(tab["N"], otab["N"][oo]), "visit counts differ at turn %d" % t
        assert np.array_equal(tab["age"], otab["age"][oo])
        np.testing.assert_allclose(tab["W"], otab["W"][oo], rtol=0, atol=1e-4)
        st = eng.search_stats()
        assert st["evals"] == agent.stat("evals") and st["node_visits"] == agent.stat("node_visits")
        assert st["subgame_tics"] == agent.stat("subgame_tics") and st["subgames"] == agent.stat("subgames")
        assert st["collisions"] == 0 and st["overflow"] == 0 and agent.stat("alias_errors") == 0
        # root tic on both sides (native spawn) with the GPU's root moves
        actions = np.ones((G, 8), np.uint8)
        for (g_i, s), m in zip(ids, root_moves):
            actions[g_i, s] = m
        eng.step(actions=torch.from_numpy(actions).cuda(), spawn_mode=2, tic=True, encode=False)
        r0 = 0
        for g_i in range(G):
            if done[g_i]:
                continue
            n = games[g_i].n_live
            if games[g_i].tic(root_moves[r0:r0 + n].astype(np.int32), spawn_mode=2, chance=0.15, seed=seed):
                done[g_i] = True
            r0 += n
        for g_i in range(0, G, 5):
            assert_dump_equal(eng.get_state(g_i), games[g_i].dump(), "root game %d after turn %d" % (g_i, t))
    eng.close()
