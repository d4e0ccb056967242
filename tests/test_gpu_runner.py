"""GPU: the drop-in classes (MPGameRunner / Agent / AlphaNNet / pit runner) keep the reference's surface and agree
with the CPU oracle."""
import numpy as np
import pytest

from tests.helpers import assert_dump_equal

pytestmark = pytest.mark.gpu


class RandomAgent:
    """any object with make_moves(games, ids) works as Alice (mp_game_runner.py:43)"""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.log = []

    def make_moves(self, games, ids):
        mv = self.rng.integers(0, 3, size=len(ids)).tolist()
        self.log.append((list(ids), mv))
        return mv


def test_runner_with_host_agent_matches_oracle():
    from alphasnake_zero_b200.utils.mp_game_runner import MPGameRunner
    from oracle import oracle as orc
    G, seed = 48, 21
    gr = MPGameRunner(11, 11, 4, 1, G, seed=seed, verbose=False)
    alice = RandomAgent(3)
    rewards = gr.run(alice)
    assert len(rewards) == G and all(r is not None and len(r) == 4 for r in rewards)
    games = []
    for gi in range(G):
        g = orc.OracleGame(11, 11, 4, 1); g.init_native(seed, gi, 0); games.append(g)
    done = [False] * G
    want = [None] * G
    for ids, mv in alice.log:
        per = {}
        for (g, s), m in zip(ids, mv):
            per.setdefault(g, []).append(m)
        assert sorted(per.keys()) == [g for g in range(G) if not done[g]]
        for g, ms in per.items():
            assert [s for (gg, s) in ids if gg == g] == games[g].live_ids()     # ids order = live-list order
            if games[g].tic(np.array(ms, np.int32), spawn_mode=2, chance=0.15, seed=seed):
                done[g] = True
                want[g] = [None if r == 0 else float(r) for r in games[g].dump()["snake"][:, 5]]
    assert all(done)
    assert rewards == want
    tot = np.zeros(6)
    for g in games:
        tot += g.dump()["counters"][:6]
    got = [gr.wall_collision, gr.body_collision, gr.head_collision, gr.starvation, gr.food_eaten, gr.game_length]
    np.testing.assert_allclose(got, tot / G)
    for r in rewards:
        assert sum(1 for x in r if x == 1.0) <= 1 and all(x in (None, 1.0, -1.0) for x in r)


def test_selfplay_with_search_agent_records():
    from alphasnake_zero_b200.utils.agent import Agent, StubNet
    from alphasnake_zero_b200.utils.mp_game_runner import MPGameRunner
    G = 16
    alice = Agent(StubNet(), 2, True, 8, 16)
    gr = MPGameRunner(11, 11, 4, 1, G, seed=5, verbose=False, table_log2=18)
    rewards = gr.run(alice)
    assert len(alice.records) == len(alice.values) > G * 4
    assert alice.records[0].shape == (21, 21, 3) and alice.values[0].shape == (3,)
    assert all(r is not None for r in rewards)
    assert gr.game_length > 1
    # every record is a legal plane: own head at the centre, values within the reference's ranges
    for p in alice.records[:50]:
        assert np.all(p[10, 10] == -1.0)
    alice.clear()
    assert alice.records == [] and gr.engine.table()["keys"].shape[0] == 0


def test_alphannet_torch_forward_matches_oracle():
    import torch
    from alphasnake_zero_b200.utils.alpha_nnet import AlphaNNet
    from oracle import net_oracle as no
    from oracle import oracle as orc
    w = no.init_weights(11, seed=3, randomize_bn=True)
    net = AlphaNNet(weights=w, backend="torch", dtype="fp32")
    g = orc.OracleGame(); g.init_native(1, 0)
    X = []
    rng = np.random.default_rng(0)
    for t in range(6):
        X += g.get_states()
        g.tic(rng.integers(0, 3, g.n_live).astype(np.int32), spawn_mode=2, seed=1)
    X = np.array(X[:12], np.float32)
    want = no.forward(w, X)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    got = net.forward_torch(torch.from_numpy(X).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-5)     # fp32 tolerance of BASELINE.json's north star
    np.testing.assert_array_equal(net.v(X) == -1.0, no.v(w, X) == -1.0)
    assert net.v(X).shape == (12, 3) and net.v(X).dtype == np.float32


def test_pit_runner_and_pit_agent():
    from alphasnake_zero_b200.utils.alpha_nnet import AlphaNNet
    from alphasnake_zero_b200.utils.pit_agent import Agent as PitAgent
    from alphasnake_zero_b200.utils.pit_mp_game_runner import MPGameRunner as PitRunner
    a = AlphaNNet(input_shape=(21, 21, 3), seed=1, backend="torch", dtype="fp32")
    b = AlphaNNet(input_shape=(21, 21, 3), seed=2, backend="torch", dtype="fp32")
    gr = PitRunner(11, 11, 2, 1, 24, seed=4)
    winners = gr.run(PitAgent(a), PitAgent(b), 1)
    assert len(winners) == 24 and all(w in (None, 0, 1) for w in winners)
    gr = PitRunner(11, 11, 4, 1, 8, seed=4)
    winners = gr.run(PitAgent(a), PitAgent(b))
    assert all(w in (None, 0, 1, 2, 3) for w in winners)


@pytest.mark.parametrize("name", ["1v1", "2v2", "1v3", "3v1_7x7"])
def test_pit_runner_against_reference(name):
    """pit_mp_game_runner.MPGameRunner.run(Alice, Bob, Alice_snake_cnt) of the unmodified reference, recorded with two
    stand-in value functions (tests/golden/make_golden.py gen_pit): same start layouts, the reference's food cells replayed;
    every move of every turn and the winner list -- team split (:30-34), winner of an ended game (:44-48) and the early
    exit when one team is gone (:49-60) -- must be identical."""
    from alphasnake_zero_b200.utils.pit_agent import Agent as PitAgent
    from alphasnake_zero_b200.utils.pit_mp_game_runner import MPGameRunner as PitRunner
    from tests.helpers import KeyStubNet, load
    from tests.test_gpu_env import init_dump
    z = load("pit_%s.npz" % name)
    side, S, dec, G = int(z["H"]), int(z["S"]), int(z["health_dec"]), int(z["G"])
    acnt = None if int(z["alice_cnt"]) < 0 else int(z["alice_cnt"])
    gr = PitRunner(side, side, S, dec, G, seed=1)
    for gi in range(G):
        nf = int(z["init_nfood"][gi])
        gr.engine.set_state(gi, init_dump(side, S, z["init_start"][gi], z["init_last"][gi], z["init_food"][gi][:nf]))
    log = []
    winners = gr.run(PitAgent(KeyStubNet(1)), PitAgent(KeyStubNet(0)), acnt, spawn_trace=z["spawn"], move_log=log)
    assert [-1 if w is None else w for w in winners] == z["winners"].tolist()
    assert len(log) == z["moves"].shape[0]
    for t, played in enumerate(log):
        assert np.array_equal(played, z["moves"][t]), "moves differ in turn %d" % t


def test_game_view_surface():
    from alphasnake_zero_b200.utils.mp_game_runner import MPGameRunner
    from oracle import oracle as orc
    gr = MPGameRunner(11, 11, 4, 1, 3, seed=9, verbose=False)
    gr._make_engine(RandomAgent(0))
    g = gr.games[1]
    og = orc.OracleGame(); og.init_native(9, 1, 0)
    assert g.get_ids() == [(1, s) for s in range(4)]
    st = g.get_states()
    for k in range(4):
        assert np.array_equal(st[k].view(np.uint32), og.make_state(k).view(np.uint32))
    assert len(g.snakes) == 4 and g.snakes[0].length == 3 and g.rewards == [None] * 4
    assert len(g.food) >= 2


def test_game_seam_ctor_tic_subgame():
    """SURVEY 8(b) last row: Game(ID, height, width, snake_cnt, health_dec, food_spawn_chance), .tic(moves), .subgame(ID)
    (game.py:13, 87, 266) for callers that step one game by hand -- standalone games and views of a runner's games"""
    from alphasnake_zero_b200.utils.game import Game
    from alphasnake_zero_b200.utils.mp_game_runner import MPGameRunner
    from oracle import oracle as orc
    from tests.helpers import assert_dump_equal
    rng = np.random.default_rng(3)
    g = Game(7, 11, 11, 4, 1, 0.15)
    og = orc.OracleGame(11, 11, 4, 1); og.init_native(7, 0, 0)
    assert g.height == 11 and g.snake_cnt == 4 and g.rewards == [None] * 4 and len(g.snakes) == 4
    sub_checked = False
    for t in range(200):
        mv = rng.integers(0, 3, size=og.n_live).tolist()
        if t == 3:
            sg, osg = g.subgame(99), og.clone()               # never spawns food, counters restart, state copied
            assert sg.id == 99 and sg.food_spawn_chance == 0.0 and sg.food == g.food
            for _ in range(4):
                m2 = rng.integers(0, 3, size=osg.n_live).tolist()
                r2 = sg.tic(m2)
                e2 = osg.tic(np.array(m2, np.int32), spawn_mode=0)
                got, want = sg._dump(), osg.dump()
                for k in ("owner", "dist", "food"):
                    assert np.array_equal(got[k], want[k])
                assert np.array_equal(got["snake"][:, :3], want["snake"][:, :3])
                assert (r2 != 0) == bool(e2)
                if e2:
                    break
            assert_dump_equal(g._dump(), og.dump(), "the root game is untouched by its sub-game")
            sub_checked = True
        res = g.tic(mv)
        ended = og.tic(np.array(mv, np.int32), spawn_mode=2, chance=0.15, seed=7)
        assert_dump_equal(g._dump(), og.dump(), "standalone game tic %d" % t)
        if ended:
            want = [None if r == 0 else float(r) for r in og.dump()["snake"][:, 5]]
            assert res == want and res.count(1.0) <= 1
            break
        assert res == 0
    assert sub_checked and ended
    with pytest.raises(ValueError):
        Game(1, 11, 11, 4).tic([1, 1])                        # one move per live snake
    # a view of a runner's game: tic steps that game only
    gr = MPGameRunner(11, 11, 4, 1, 5, seed=9, verbose=False)
    gr._make_engine(RandomAgent(0))
    before = [gr.engine.get_state(i) for i in range(5)]
    assert gr.games[2].tic([1, 1, 1, 1]) == 0
    for i in (0, 1, 3, 4):
        assert_dump_equal(gr.engine.get_state(i), before[i], "game %d must not move" % i)
    o2 = orc.OracleGame(11, 11, 4, 1); o2.load_dump(before[2])
    o2.tic(np.array([1, 1, 1, 1], np.int32), spawn_mode=0)
    got = gr.engine.get_state(2)
    assert np.array_equal(got["snake"][:, :5], o2.dump()["snake"][:, :5]) and got["counters"][5] == 1


def test_single_game_writes_replay_rep(tmp_path, monkeypatch):
    """test_model.py:13-23: one game with the search agent, replay.rep gets two frames per tic (game.py:140-141,194-195) in
    the text format player.py:63-79 parses"""
    import ast
    from alphasnake_zero_b200.utils.agent import Agent, StubNet
    from alphasnake_zero_b200.utils.mp_game_runner import MPGameRunner
    monkeypatch.chdir(tmp_path)
    gr = MPGameRunner(11, 11, 4, 9, 1, verbose=False, seed=4)
    rewards = gr.run(Agent(StubNet(), 100, False, 4, 8))
    assert len(rewards) == 1 and rewards[0].count(1.0) <= 1
    pages = [p for p in open("replay.rep").read().split("\n\n") if p.strip()]
    assert len(pages) == 2 * int(round(gr.game_length)) > 4
    for page in pages:
        rows = [ast.literal_eval(line) for line in page.strip().split("\n")]
        assert len(rows) == 11 and all(len(r) == 11 and all(-4 <= v <= 9 for v in r) for r in rows)
    first = [ast.literal_eval(line) for line in pages[0].strip().split("\n")]
    assert sum(v < 0 for r in first for v in r) >= 1 and any(v == 9 for r in first for v in r)
