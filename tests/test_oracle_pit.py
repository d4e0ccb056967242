"""CPU: the pit fixtures (reference pit_mp_game_runner.MPGameRunner.run with two stand-in value functions) replayed over the
oracle's games: every move, and the winner list including the early exit (pit_mp_game_runner.py:49-60)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import KeyStubNet, load

PITS = ["1v1", "2v2", "1v3", "3v1_7x7"]


def argmaxs(V):          # pit_agent.py:15-28
    return [(0 if v[0] > v[2] else 2) if v[0] > v[1] else (1 if v[1] > v[2] else 2) for v in V]


@pytest.mark.parametrize("name", PITS)
def test_pit_replay(name):
    z = load("pit_%s.npz" % name)
    side, S, dec, G = int(z["H"]), int(z["S"]), int(z["health_dec"]), int(z["G"])
    acnt = S // 2 if int(z["alice_cnt"]) < 0 else int(z["alice_cnt"])
    nets = (KeyStubNet(1), KeyStubNet(0))
    games = {}
    for gi in range(G):
        g = orc.OracleGame(side, side, S, dec)
        nf = int(z["init_nfood"][gi])
        g.init_explicit(z["init_start"][gi], z["init_last"][gi], z["init_food"][gi][:nf])
        games[gi] = g
    winners = [-1] * G
    turn = 0
    while games:
        for gi in list(games):
            g = games[gi]
            live = g.live_ids()
            mv = []
            for k, sid in enumerate(live):
                m = argmaxs(nets[0 if sid < acnt else 1].v(g.make_state(k)[None]))[0]
                assert m == int(z["moves"][turn, gi, sid]), (name, turn, gi, sid)
                mv.append(m)
            assert int(z["spawn"][turn, gi]) != -2
            ended = g.tic(np.array(mv, np.int32), spawn_mode=1, spawn_cell=int(z["spawn"][turn, gi]))
            d = g.dump()["snake"]
            if ended:
                w = np.nonzero(d[:, 5] == 1)[0]
                winners[gi] = int(w[0]) if len(w) else -1
                del games[gi]
            else:
                live = g.live_ids()
                if all(s < acnt for s in live) or all(s >= acnt for s in live):
                    winners[gi] = live[0]
                    del games[gi]
        turn += 1
    assert winners == z["winners"].tolist()
    assert turn == z["spawn"].shape[0]
