"""GPU parity of the search kernels (through the C ABI):
 (1) replaying the in-tree moves the unmodified reference sampled: visit counts, key sets and ages bit-exact,
     Q/W within float32 reassociation (the reference backs up sequentially, the GPU with atomics);
 (2) native GPU sampling recorded as a trace, replayed by the CPU oracle."""
import numpy as np
import pytest

from tests.helpers import assert_dump_equal, load
from tests.test_gpu_env import _engine, init_dump

pytestmark = pytest.mark.gpu

# the last three (round 2) start from hand-built / mid-game states: sub-games in which a head-on winner lives on with
# health <= 0 at health_dec 9 and 3 (game.py:156-165), and 19x19x8 with 3-4 live snakes (depth 4-6, deaths, evictions)
MCTS = ["11x11x4_train", "11x11x4_eval", "7x7x4_dec9", "19x19x8", "11x11x4_b5",
        "11x11x4_neghealth_dec9", "11x11x4_neghealth_dec3", "19x19x8_mid"]


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
    custom = "custom" in z.files and int(z["custom"]) == 1
    for gi in range(G):
        if custom:      # hand-built or mid-game start: the state before the first recorded root turn
            eng.set_state(gi, {k: z["t0_before_%s" % k][gi] for k in ("snake", "owner", "dist", "food", "counters")})
        else:
            nf = int(z["init_nfood"][gi])
            eng.set_state(gi, init_dump(side, S, z["init_start"][gi], z["init_last"][gi], z["init_food"][gi][:nf]))
    if "nonpositive_health_tics" in z.files and "neghealth" in name:
        assert int(z["nonpositive_health_tics"]) > 0
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
    (11, 2, 40, 6, 8, 10.0, True, 12), (11, 4, 768, 8, 32, 2.0, True, 2)])
def test_native_search_against_oracle(side, S, G, D, breadth, base, training, turns):
    """The GPU samples with its own RNG and records the trace; the oracle replays it."""
    _native_search_against_oracle(side, S, G, D, breadth, base, training, turns, 20)


def test_native_search_across_table_compactions():
    """a table of 2^12 slots: the live entries are compacted into a fresh table several times during the run (more than
    half of the slots taken, many of them by evicted entries, agent.py:101-110); keys, visit counts, W and ages still match"""
    n = _native_search_against_oracle(11, 4, 32, 4, 16, 2.0, True, 12, 12)
    assert n >= 2, "the run did not compact the table (%d times)" % n


def test_native_search_with_compaction_inside_a_turn():
    """Evicted entries keep their slots until a compaction.  With per-step host synchronisation (the network path) the engine
    sees the occupancy and compacts between two epochs of one root turn when the table is more than 3/4 full; the result
    is still the oracle's, entry for entry."""
    n = _native_search_against_oracle(11, 4, 32, 4, 16, 2.0, True, 14, 12, sync_steps=True, want_mid=True)
    assert n >= 2


# BASELINE.json's own search shapes -- breadth 100 / depth 8 (configs[2]), breadth 200 (configs[3]) and 19x19 with
# 8 snakes at breadth 400 (configs[4]) from a mid-game start (3-5 live snakes: depth 2-6 instead of one tic per sub-game, deaths,
# evictions, table compactions) -- at the game counts the single-threaded CPU oracle replays in about half a minute each
# (768 of 4,096 games = 73,728 sub-games and about 1.2 million node visits per root turn; the per-game work is independent
# of the game count, and the full 4,096-game run is checked for overflow / collisions by bench.py's self-play leg).
@pytest.mark.parametrize("side,S,G,D,breadth,turns,table_log2,warm,compactions", [
    (11, 4, 768, 8, 100, 2, 0, 6, 0),
    (11, 4, 384, 8, 200, 2, 0, 12, 0),
    (19, 8, 16, 8, 400, 12, 20, 40, 1)])
def test_native_search_at_baseline_config_shapes(side, S, G, D, breadth, turns, table_log2, warm, compactions):
    # At these sizes the same state often is the current node of one row and an ancestor on the path of an EARLIER row of the same
    # step.  The reference's `Q_row` then aliases an entry the earlier row's backup has already updated (agent.py:180,208-220 is a
    # sequential loop), while the kernels compute every r-hat of a step from the values before the step's backups.  The oracle can
    # do either (OracleAgent.set_rhat_mode): against the order-free mode the engine must match to float reassociation at full
    # size; the distance between the oracle's two modes (visit counts, keys, ages, moves identical; a few Q values moving in the
    # fourth digit) is what tests/test_oracle_mcts.py::test_rhat_order_modes measures.
    n = _native_search_against_oracle(side, S, G, D, breadth, 2.0, True, turns, table_log2, warm_tics=warm, dump_every=97,
                                      q_atol=1e-4, rhat_before=True)   # W is a float sum of thousands of terms in a different order
    assert n >= compactions, "expected at least %d table compactions, saw %d" % (compactions, n)



def _assert_close_up_to_conditioning(got, want, atol, what):
    """Float parity at sizes where the reference's own arithmetic is ill-conditioned.  softermax raises the base to arctanh(Q)
    (agent.py:113-116) and d arctanh / dQ = 1 / (1 - Q^2): next to Q = -1 (a move that died in every visit but a few, common
    once thousands of visits share a table) one float32 ulp of Q, i.e. a different order of the same W sum, moves the pmf by
    10^-3 and every r-hat = pmf . Q above it with it.  Visit counts, keys, ages and moves stay bit-exact (asserted by the
    caller); for the float sums 99.9 % of the values must agree to `atol`, and no value may be further off than the
    amplification of a one-ulp difference allows (5e-2 is 1 ulp at |Q| = 1 - 6e-7, the closest float32 gets to 1 below it)."""
    d = np.abs(np.asarray(got, np.float64) - np.asarray(want, np.float64))
    q999, worst = float(np.quantile(d, 0.999)), float(d.max())
    assert q999 <= atol and worst < 5e-2, "%s: 99.9 %% quantile %.3g (limit %.3g), max %.3g" % (what, q999, atol, worst)


def _native_search_against_oracle(side, S, G, D, breadth, base, training, turns, table_log2, sync_steps=False, want_mid=False,
                                  warm_tics=0, dump_every=5, q_atol=5e-6, rhat_before=False):
    import os
    import torch
    from oracle import oracle as orc
    seed = 77
    eng = _engine(side=side, snakes=S, health_dec=1, games=G, seed=seed, max_depth=D, max_breadth=breadth,
                  softmax_base=base, training=training, table_log2=table_log2)
    eng.reset()
    info = eng.search_info()
    games = []
    for gi in range(G):
        g = orc.OracleGame(side, side, S, 1); g.init_native(seed, gi, 0); g.set_ids(gi, 0); games.append(g)
    if warm_tics:
        # games of every age before the search starts: uniform-random play with in-place reset, the same Philox streams on both sides
        for _ in range(warm_tics):
            eng.step(spawn_mode=2, tic=True, encode=False, auto_reset=True, random_actions=True)
        orc.env_run(G, side, side, S, 1, 0.15, seed, warm_tics, encode=False, n_threads=os.cpu_count() or 1, games=games)
        for g_i in range(0, G, dump_every):
            assert_dump_equal(eng.get_state(g_i), games[g_i].dump(), "game %d after the warm-up" % g_i)
        live_counts = [g.n_live for g in games]
        assert min(live_counts) >= 2 and len(set(live_counts)) > 1        # running games with different numbers of snakes
    agent = orc.OracleAgent(base=base, training=training, max_depth=D, max_breadth=breadth)
    agent.set_rhat_mode(rhat_before)
    done = [False] * G
    compactions, last_occupied, last_inserts = 0, 0, 0
    for t in range(turns):
        tree = torch.full((info["epochs"], info["max_steps"], G * info["P"], S), 255, dtype=torch.uint8, device="cuda")
        q, mv = eng.search(value_fn=None, trace=tree, trace_mode=2, sync_steps=sync_steps)
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
        if q_atol <= 5e-6:
            np.testing.assert_allclose(got_q, oq, rtol=0, atol=q_atol)
        else:
            _assert_close_up_to_conditioning(got_q, oq, q_atol, "root Q")
        if training:
            assert np.array_equal(omv, root_moves)
        tab, otab = eng.table(), agent.table()
        oo = np.lexsort((otab["keys"][:, 1], otab["keys"][:, 0]))
        assert np.array_equal(tab["keys"], otab["keys"][oo]), "key sets differ at turn %d" % t
        assert np.array_equal(tab["N"], otab["N"][oo]), "visit counts differ at turn %d" % t
        assert np.array_equal(tab["age"], otab["age"][oo])
        if q_atol <= 5e-6:
            np.testing.assert_allclose(tab["W"], otab["W"][oo], rtol=0, atol=1e-4)
        else:
            # float sums of up to 10^4 terms in a different order, and the rounding of a heavily visited child's Q reaches its
            # parents through r-hat: the error per visit is what is bounded (all but 0.1 % of the entries within q_atol)
            n_vis = np.maximum(tab["N"], 1.0)
            _assert_close_up_to_conditioning(tab["W"] / n_vis, otab["W"][oo] / n_vis, q_atol, "W per visit")
        st = eng.search_stats()
        assert st["evals"] == agent.stat("evals") and st["node_visits"] == agent.stat("node_visits")
        assert st["subgame_tics"] == agent.stat("subgame_tics") and st["subgames"] == agent.stat("subgames")
        assert st["collisions"] == 0 and st["overflow"] == 0 and agent.stat("alias_errors") == 0
        if st["occupied"] - last_occupied < st["inserts"] - last_inserts:
            compactions += 1            # slots were given back: the table was rebuilt from its live entries
        last_occupied, last_inserts = st["occupied"], st["inserts"]
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
        for g_i in range(0, G, dump_every):
            assert_dump_equal(eng.get_state(g_i), games[g_i].dump(), "root game %d after turn %d" % (g_i, t))
    st = eng.search_stats()
    assert st["compactions"] >= compactions, (st, compactions)
    if want_mid:
        assert st["mid_turn_compactions"] >= 1, st
    eng.close()
    return st["compactions"]


def test_policy_functions_against_reference():
    """softermax / numpy.random.choice / argmaxs kernels against known answers recorded from the reference"""
    import ctypes as C
    import torch
    from alphasnake_zero_b200 import _lib
    z = load("funcs.npz")
    Z = torch.from_numpy(z["Z"]).cuda()
    u = torch.from_numpy(z["choice_u"]).cuda()
    n = Z.shape[0]
    L = _lib.lib()
    for base in (2, 3, 10, 100):
        pmf = torch.zeros(n, 3, device="cuda"); ch = torch.zeros(n, dtype=torch.int32, device="cuda"); am = torch.zeros_like(ch)
        _lib.check(L.asz_debug_policy(C.c_void_p(Z.data_ptr()), C.c_void_p(u.data_ptr()), n, float(base), C.c_void_p(pmf.data_ptr()),
                                      C.c_void_p(ch.data_ptr()), C.c_void_p(am.data_ptr()), None))
        torch.cuda.synchronize()
        want = z["softermax_%d" % base]
        np.testing.assert_allclose(pmf.cpu().numpy(), want, rtol=4e-6, atol=1e-7)    # CUDA powf/atanhf vs NumPy: a few ulp
        assert np.array_equal(am.cpu().numpy(), z["argmaxs"])
        if base == 2:
            # the draw is a function of (pmf, u); it may only differ where u sits within rounding of a cdf boundary
            cdf = np.cumsum(want.astype(np.float64), axis=1); cdf /= cdf[:, -1:]
            safe = np.abs(cdf[:, :2] - z["choice_u"][:, None]).min(axis=1) > 1e-5
            assert np.array_equal(ch.cpu().numpy()[safe], z["choice_idx"][safe]) and safe.sum() > 200


def test_obstacle_mask_kernel():
    import ctypes as C
    import torch
    from oracle import oracle as orc
    from alphasnake_zero_b200 import _lib
    from tests.test_gpu_net import game_planes
    X = game_planes(11, 4, 200, seed=5)
    for numpy1 in (False, True):
        eng = _engine(side=11, snakes=4, games=1, numpy1_mask=numpy1)
        v = torch.full((len(X), 3), 0.25, device="cuda")
        xs = torch.from_numpy(X).cuda()
        _lib.check(eng.L.asz_obstacle_mask(eng.h, C.c_void_p(xs.data_ptr()), len(X), C.c_void_p(v.data_ptr()), None))
        got = v.cpu().numpy()
        c = 10
        for i in range(len(X)):
            b = [X[i, c, c - 1, 1], X[i, c - 1, c, 1], X[i, c, c + 1, 1]]
            want = [(-1.0 if ((float(x) >= 0.04) if numpy1 else (x >= np.float32(0.04))) else 0.25) for x in b]
            assert got[i].tolist() == want
            if not numpy1:
                assert np.array_equal(got[i], orc.obstacle_mask(X[i], 11, 11, np.full(3, 0.25, np.float32)))
        eng.close()
    # the two semantics differ exactly on dist == 2 cells
    assert np.float32(2 * 0.02) >= np.float32(0.04) and not (float(np.float32(2 * 0.02)) >= 0.04)


def test_search_with_network_value_function_against_oracle():
    """the full Agent.make_moves path with a real value network: the engine calls the network on its eval batch, the
    oracle calls the same network through its value-function hook; moves are recorded by the GPU and replayed."""
    import torch
    from alphasnake_zero_b200.utils.alpha_nnet import AlphaNNet
    from oracle import oracle as orc
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = AlphaNNet(input_shape=(21, 21, 3), seed=4, backend="torch", dtype="fp32")
    for k in ("dense2_w",):
        net.weights[k] = (net.weights[k] * 5).astype(np.float32)       # spread the values (|v| stays well below 1: Q == 1 is undefined in the reference)

    def vf_device(planes):
        return net.forward_torch(planes, torch.float32)

    def vf_host(planes):        # AlphaNNet.v contract: forward + obstacle mask
        return net.v(planes)
    G, S, D, B, seed = 6, 4, 8, 16, 31
    eng = _engine(side=11, snakes=S, games=G, seed=seed, max_depth=D, max_breadth=B, softmax_base=2.0, training=True, table_log2=18)
    eng.reset()
    info = eng.search_info()
    games = []
    for gi in range(G):
        g = orc.OracleGame(11, 11, S, 1); g.init_native(seed, gi, 0); g.set_ids(gi, 0); games.append(g)
    agent = orc.OracleAgent(base=2.0, training=True, max_depth=D, max_breadth=B, value_fn=vf_host)
    for t in range(2):
        tree = torch.full((info["epochs"], info["max_steps"], G * info["P"], S), 255, dtype=torch.uint8, device="cuda")
        q, mv = eng.search(value_fn=vf_device, trace=tree, trace_mode=2)
        mvh = mv.cpu().numpy()
        ids = [(gi, s) for gi in range(G) for s in games[gi].live_ids()]
        root = np.array([mvh[g, s] for g, s in ids], np.uint8)
        omv, oq = agent.make_moves(games, G, root_turn=t, tree_moves=np.ascontiguousarray(tree.cpu().numpy()), root_moves=root.copy(), replay=True)
        tab, otab = eng.table(), agent.table()
        oo = np.lexsort((otab["keys"][:, 1], otab["keys"][:, 0]))
        assert np.array_equal(tab["keys"], otab["keys"][oo]) and np.array_equal(tab["N"], otab["N"][oo])
        np.testing.assert_allclose(tab["W"], otab["W"][oo], rtol=0, atol=2e-4)
        got_q = np.array([q.cpu().numpy()[g, s] for g, s in ids])
        np.testing.assert_allclose(got_q, oq, rtol=0, atol=2e-5)
        assert np.abs(oq[oq > -1]).max() > 0.01
        actions = np.ones((G, 8), np.uint8)
        for (g, s), m in zip(ids, root):
            actions[g, s] = m
        eng.step(actions=torch.from_numpy(actions).cuda(), spawn_mode=2, tic=True, encode=False)
        r0 = 0
        for gi in range(G):
            n = games[gi].n_live
            games[gi].tic(root[r0:r0 + n].astype(np.int32), spawn_mode=2, chance=0.15, seed=seed)
            r0 += n
    eng.close()


def test_native_whole_turn_search_equals_python_driven_loop():
    """asz_search_run_net (the product path: one native call per root turn) against the Python-driven step loop with the same
    hand-written network as value function.  The native run records its sampled moves, the Python-driven run replays them
    (W is summed with float atomics, so a 1-ulp difference could otherwise flip a draw): identical visit counts, keys,
    evaluation counts; Q within float reassociation."""
    import torch
    from alphasnake_zero_b200.utils.alpha_nnet import AlphaNNet
    net = AlphaNNet(input_shape=(21, 21, 3), seed=4, backend="native")
    net.weights["dense2_w"] = (net.weights["dense2_w"] * 5).astype(np.float32)
    nat = net._get_native()
    G, S, D, B, seed = 48, 4, 8, 24, 5
    ea = _engine(side=11, snakes=S, games=G, seed=seed, max_depth=D, max_breadth=B, softmax_base=2.0, training=True, table_log2=20)
    eb = _engine(side=11, snakes=S, games=G, seed=seed, max_depth=D, max_breadth=B, softmax_base=2.0, training=True, table_log2=20)
    ea.reset(); eb.reset()
    info = ea.search_info()
    for t in range(3):
        tree = torch.full((info["epochs"], info["max_steps"], G * info["P"], S), 255, dtype=torch.uint8, device="cuda")
        qa, ma = ea.search(net=nat, trace=tree, trace_mode=2)
        qa, ma = qa.clone(), ma.clone()
        qb, mb = eb.search(value_fn=nat.forward, trace=tree, trace_mode=1, root_trace=ma.reshape(-1).contiguous())
        assert torch.equal(ma, mb)
        np.testing.assert_allclose(qa.cpu().numpy(), qb.cpu().numpy(), rtol=0, atol=2e-5)
        ta, tb = ea.table(), eb.table()
        assert np.array_equal(ta["keys"], tb["keys"]) and np.array_equal(ta["N"], tb["N"])
        np.testing.assert_allclose(ta["W"], tb["W"], rtol=0, atol=2e-4)
        act = torch.where(ma < 3, ma, torch.ones_like(ma))
        ea.step(actions=act, spawn_mode=2, tic=True, encode=False)
        eb.step(actions=act, spawn_mode=2, tic=True, encode=False)
    sa, sb = ea.search_stats(), eb.search_stats()
    assert sa["evals"] == sb["evals"] > 0 and sa["node_visits"] == sb["node_visits"] and sa["subgames"] == sb["subgames"]
    ea.close(); eb.close()
