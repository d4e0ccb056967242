"""CPU: the search oracle (oa_make_moves) replaying the in-tree moves the reference sampled, against the
reference's own tables (N exact; Q/W within float32 rounding of powf/atanhf/dot)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import assert_dump_equal, load

MCTS = ["11x11x4_train", "11x11x4_eval", "7x7x4_dec9", "19x19x8", "11x11x4_b5",
        "11x11x4_neghealth_dec9", "11x11x4_neghealth_dec3", "19x19x8_mid"]


def replay_reference_turns(z, make_agent, on_turn):
    """Drives root games from the fixture; calls on_turn(t, games(list of OracleGame, live), z) each root turn."""
    H, W, S, dec, G = (int(z[k]) for k in ("H", "W", "S", "health_dec", "G"))
    games = {}
    custom = "custom" in z.files and int(z["custom"]) == 1
    for gi in range(G):
        g = orc.OracleGame(H, W, S, dec)
        if custom:      # hand-built or mid-game start: the state before the first recorded root turn
            g.load_dump({k: z["t0_before_%s" % k][gi] for k in ("snake", "owner", "dist", "food", "counters")})
        else:
            nf = int(z["init_nfood"][gi])
            g.init_explicit(z["init_start"][gi], z["init_last"][gi], z["init_food"][gi][:nf])
        g.set_ids(gi, 0)
        games[gi] = g
    for t in range(int(z["n_turns"])):
        live = z["t%d_live" % t].tolist()
        assert sorted(games.keys()) == live
        for j, gid in enumerate(live):
            want = {k: z["t%d_before_%s" % (t, k)][j] for k in ("snake", "owner", "dist", "food", "counters")}
            assert_dump_equal(games[gid].dump(), want, "root game %d turn %d" % (gid, t))
        moves = on_turn(t, [games[g] for g in live])
        # root tic with the reference's spawn cells
        r0 = 0
        for j, gid in enumerate(live):
            n = games[gid].n_live
            ended = games[gid].tic(moves[r0:r0 + n], spawn_mode=1, spawn_cell=int(z["t%d_spawn" % t][j]))
            r0 += n
            if ended:
                del games[gid]


@pytest.mark.parametrize("name", MCTS)
def test_mcts_replay(name):
    z = load("mcts_%s.npz" % name)
    G = int(z["G"])
    agent = orc.OracleAgent(base=float(z["base"]), training=bool(z["training"]), max_depth=int(z["D"]),
                            max_breadth=int(z["breadth"]))
    stats = dict(max_q=0.0)

    def on_turn(t, games):
        tree = np.ascontiguousarray(z["t%d_tree" % t])
        want_moves = z["t%d_root_moves" % t]
        moves, q = agent.make_moves(games, G, root_turn=t, tree_moves=tree,
                                    root_moves=np.ascontiguousarray(want_moves), replay=True)
        assert len(moves) == len(want_moves)
        np.testing.assert_allclose(q, z["t%d_root_q" % t], rtol=0, atol=2e-6)
        if bool(z["training"]):
            assert np.array_equal(moves, want_moves)
        else:
            # argmaxs: exact unless two Q values are within rounding of each other
            qq = z["t%d_root_q" % t]
            srt = np.sort(qq, axis=1)
            safe = (srt[:, 2] - srt[:, 1]) > 1e-5
            assert np.array_equal(moves[safe], want_moves[safe])
            moves = want_moves.astype(np.int32)
        tab = agent.table()
        order = np.lexsort((tab["keys"][:, 1], tab["keys"][:, 0]))
        assert np.array_equal(tab["keys"][order], z["t%d_tab_keys" % t]), "key sets differ at turn %d" % t
        assert np.array_equal(tab["N"][order], z["t%d_tab_N" % t]), "visit counts differ at turn %d" % t
        assert np.array_equal(tab["age"][order], z["t%d_tab_age" % t])
        np.testing.assert_allclose(tab["W"][order], z["t%d_tab_W" % t], rtol=0, atol=2e-5)
        np.testing.assert_allclose(tab["Q"][order], z["t%d_tab_Q" % t], rtol=0, atol=2e-6)
        stats["max_q"] = max(stats["max_q"], float(np.abs(tab["Q"][order] - z["t%d_tab_Q" % t]).max()))
        assert agent.stat("evals") == int(z["t%d_n_evals" % t])
        return moves.astype(np.int32)

    replay_reference_turns(z, None, on_turn)
    assert agent.stat("alias_errors") == 0
    if bool(z["training"]):
        planes, q = agent.records(int(z["H"]), int(z["W"]))
        assert np.array_equal(planes.view(np.uint32), z["records"].view(np.uint32))


def test_native_sampling_is_deterministic_and_recorded():
    def run():
        games = []
        for gi in range(3):
            g = orc.OracleGame(); g.init_native(3, gi); g.set_ids(gi, 0); games.append(g)
        a = orc.OracleAgent(base=2, training=True, max_depth=8, max_breadth=16)
        tree = np.full((a.epochs, 8, 3 * a.parallel, 4), 255, np.uint8)
        mv, q = a.make_moves(games, 3, root_turn=0, seed=77, tree_moves=tree)
        return mv, q, tree, a
    mv1, q1, tree1, a1 = run()
    mv2, q2, tree2, a2 = run()
    assert np.array_equal(mv1, mv2) and np.array_equal(q1, q2) and np.array_equal(tree1, tree2)
    assert (tree1 != 255).sum() == a1.stat("node_visits")
    # replaying the recorded moves reproduces the same tables bit for bit
    games = []
    for gi in range(3):
        g = orc.OracleGame(); g.init_native(3, gi); g.set_ids(gi, 0); games.append(g)
    a3 = orc.OracleAgent(base=2, training=True, max_depth=8, max_breadth=16)
    mv3, q3 = a3.make_moves(games, 3, root_turn=0, tree_moves=tree1, root_moves=mv1.astype(np.uint8), replay=True)
    assert np.array_equal(mv3, mv1) and np.array_equal(q3.view(np.uint32), q1.view(np.uint32))
    t1, t3 = a1.table(), a3.table()
    o1 = np.lexsort((t1["keys"][:, 1], t1["keys"][:, 0])); o3 = np.lexsort((t3["keys"][:, 1], t3["keys"][:, 0]))
    for k in ("keys", "Q", "W", "N", "age"):
        assert np.array_equal(t1[k][o1], t3[k][o3])


def test_rhat_order_modes():
    """The one place where a parallel search cannot follow the reference's sequential loop to the letter (agent.py:208-220): r-hat of
    a row is computed from a LIVE alias of its table entry, so it sees the backups of earlier rows of the same step when its state is
    one of their ancestors.  The oracle implements both orders.  Same sampled moves => identical keys, visit counts and ages; W / Q
    differ for a small fraction of the entries, in the fourth digit at most (this is the documented deviation of the CUDA search)."""
    def run(before):
        games = []
        for gi in range(24):
            g = orc.OracleGame(); g.init_native(5, gi); g.set_ids(gi, 0); games.append(g)
        a = orc.OracleAgent(base=2, training=True, max_depth=8, max_breadth=32)
        a.set_rhat_mode(before)
        tree = run.tree if before else np.full((a.epochs, 8, 24 * a.parallel, 4), 255, np.uint8)
        if before:
            mv, q = a.make_moves(games, 24, root_turn=0, tree_moves=tree, root_moves=run.mv.astype(np.uint8), replay=True)
        else:
            mv, q = a.make_moves(games, 24, root_turn=0, seed=9, tree_moves=tree)
            run.tree, run.mv = tree, mv
        t = a.table()
        o = np.lexsort((t["keys"][:, 1], t["keys"][:, 0]))
        return q, {k: v[o] for k, v in t.items()}
    q0, t0 = run(False)
    q1, t1 = run(True)
    assert np.array_equal(t0["keys"], t1["keys"]) and np.array_equal(t0["N"], t1["N"]) and np.array_equal(t0["age"], t1["age"])
    dq = np.abs(t0["Q"] - t1["Q"])
    assert dq.max() < 5e-3 and np.mean(dq > 1e-5) < 0.05, (dq.max(), np.mean(dq > 1e-5))
    assert np.abs(q0 - q1).max() < 5e-3
