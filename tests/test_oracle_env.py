"""CPU: the oracle (oracle/asz_oracle.c) against fixtures produced by the unmodified reference."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import assert_dump_equal, digest, golden_dump, load

ENVS = ["11x11x4", "11x11x4_dec9", "7x7x4", "19x19x8", "11x11x2", "7x7x8"]


@pytest.mark.parametrize("name", ENVS)
def test_env_replay(name):
    z = load("env_%s.npz" % name)
    H, W, S, dec = int(z["H"]), int(z["W"]), int(z["S"]), int(z["health_dec"])
    gp = z["game_ptr"]
    full = {(int(t), int(k)): i for i, (t, k) in enumerate(z["full_idx"])}
    n_planes = 0
    for gi in range(len(gp) - 1):
        g = orc.OracleGame(H, W, S, dec)
        nf = int(z["init_nfood"][gi])
        g.init_explicit(z["init_start"][gi], z["init_last"][gi], z["init_food"][gi][:nf])
        for t in range(int(gp[gi]), int(gp[gi + 1])):
            n = int(z["nlive"][t])
            assert g.n_live == n
            ended = g.tic(z["moves"][t][:n].astype(np.int32), spawn_mode=1, spawn_cell=int(z["spawn"][t]))
            assert ended == int(z["ended"][t]), (name, gi, t)
            assert_dump_equal(g.dump(), golden_dump(z, t), "%s game %d tic %d" % (name, gi, t))
            d0, d1 = int(z["dig_ptr"][t]), int(z["dig_ptr"][t + 1])
            assert g.n_live == d1 - d0 or (ended and d1 - d0 == g.n_live)
            for k in range(d1 - d0):
                p = g.make_state(k)
                assert digest(p) == z["digests"][d0 + k], (name, gi, t, k)
                if (t, k) in full:
                    assert np.array_equal(p.view(np.uint32), z["full_planes"][full[(t, k)]].view(np.uint32))
                n_planes += 1
    assert n_planes == len(z["digests"])


def test_edge_cases():
    z = load("edge_cases.npz")
    for i, name in enumerate(z["names"]):
        g = orc.OracleGame(11, 11, 4, int(z["health_dec"][i]))
        before = {k: z["before_" + k][i] for k in ("snake", "owner", "dist", "food", "counters")}
        after = {k: z["after_" + k][i] for k in ("snake", "owner", "dist", "food", "counters")}
        g.load_dump(before)
        assert_dump_equal(g.dump(), before, "load/dump round trip %s" % name)
        pb = z["planes_before"][int(z["pb_ptr"][i]):int(z["pb_ptr"][i + 1])]
        for k in range(len(pb)):
            assert np.array_equal(g.make_state(k).view(np.uint32), pb[k].view(np.uint32)), name
        n = g.n_live
        ended = g.tic(z["moves"][i][:n].astype(np.int32), spawn_mode=0)
        assert ended == int(z["ended"][i]), name
        assert_dump_equal(g.dump(), after, str(name))
        pa = z["planes_after"][int(z["pa_ptr"][i]):int(z["pa_ptr"][i + 1])]
        assert len(pa) == g.n_live, name
        for k in range(len(pa)):
            assert np.array_equal(g.make_state(k).view(np.uint32), pa[k].view(np.uint32)), name


def test_edge_sequences():
    """multi-tic scenarios recorded from the reference: a head-on winner ends a tic alive with health <= 0 (game.py:156-165 is an
    elif chain), then starves / eats / hits the wall / wins again; health_dec 9, 3 and 1"""
    z = load("edge_sequences.npz")
    keys = ("snake", "owner", "dist", "food", "counters")
    assert int(z["min_health"].min()) < 0
    for i, name in enumerate(z["names"]):
        g = orc.OracleGame(11, 11, 4, int(z["health_dec"][i]))
        g.load_dump({k: z["before_" + k][i] for k in keys})
        for t in range(int(z["tic_ptr"][i]), int(z["tic_ptr"][i + 1])):
            n = g.n_live
            ended = g.tic(z["moves"][t][:n].astype(np.int32), spawn_mode=0)
            assert ended == int(z["ended"][t]), (name, t)
            assert_dump_equal(g.dump(), {k: z["after_" + k][t] for k in keys}, "%s tic %d" % (name, t))
            pl = z["planes"][int(z["pl_ptr"][t]):int(z["pl_ptr"][t + 1])]
            assert len(pl) == (0 if ended else g.n_live), name
            for k in range(len(pl)):
                assert np.array_equal(g.make_state(k).view(np.uint32), pl[k].view(np.uint32)), (name, t, k)


def test_food_chance_zero_never_spawns():
    """game.py:130 `if self.food_spawn_chance > 0.0`: with chance 0 no food is ever spawned, not even on a board without food"""
    g = orc.OracleGame(11, 11, 2, 1)
    g.init_explicit([(1, 1), (9, 9)], [1, 3], [])
    for _ in range(6):
        g.tic(np.array([1, 1], np.int32), spawn_mode=2, chance=0.0, seed=3)
        assert g.dump()["food"].sum() == 0
    g.tic(np.array([2, 2], np.int32), spawn_mode=2, chance=1e-12, seed=3)    # any positive chance spawns on a board without food
    assert g.dump()["food"].sum() == 1


def test_funcs():
    z = load("funcs.npz")
    Z = z["Z"]
    for base in (2, 3, 10, 100):
        want = z["softermax_%d" % base]
        got = np.array([orc.softermax(zz, base) for zz in Z])
        # libm powf/atanhf vs NumPy's float32 loops: a few ulp
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7)
        assert np.array_equal(got[0], np.full(3, np.float32(1.0 / 3.0)))   # all masked -> uniform
        some = ~np.all(Z == -1.0, axis=1)
        assert np.all(got[some][Z[some] == -1.0] == 0.0)                  # masked moves are never sampled
    assert np.array_equal(np.array([orc.argmax3(zz) for zz in Z]), z["argmaxs"])
    got = np.array([orc.choice3(p, u) for p, u in zip(z["softermax_2"], z["choice_u"])])
    assert np.array_equal(got, z["choice_idx"])


def test_native_rng_and_env_batch_determinism():
    a = orc.env_run(64, tics=60, seed=5, n_threads=1)
    b = orc.env_run(64, tics=60, seed=5, n_threads=4)
    assert a == b
    assert a["steps"] == 64 * 60 and a["episodes"] > 0 and a["planes"] > a["steps"]
    c = orc.env_run(64, tics=60, seed=6, n_threads=2)
    assert c["plane_checksum"] != a["plane_checksum"]
    # native init draws a legal standard layout
    g = orc.OracleGame()
    for gid in range(50):
        g.init_native(9, gid)
        d = g.dump()
        heads = d["snake"][:, 4]
        assert len(set(heads.tolist())) == 4
        assert d["food"].sum() >= 2 and d["food"][5 * 11 + 5] == 1
        assert np.all(d["dist"][heads] == 3)
