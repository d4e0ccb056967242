"""Generate the golden fixtures in tests/golden/ by executing the UNMODIFIED reference
(/root/reference/code/utils/{game,agent,mp_game_runner}.py) in the build container.

Run:  python tests/golden/make_golden.py            (needs /root/reference; not needed on the GPU box)

The reference has no tests or golden vectors of its own, and its RNG use
(CPython `random`, set iteration order) is not portable, so every fixture records OUTCOMES
(start layouts, moves, spawned food cells, sampled tree moves) next to the resulting states.

Fixtures (all numpy .npz, compressed):
  env_<name>.npz    random-play games: initial layouts, per-tic moves / spawn cell / post-tic canonical
                    dump / blake2b-64 digest of every live snake's plane, plus full planes of a few tics
  edge_cases.npz    hand-built single-tic scenarios for every trap of SURVEY.md Appendix D
  edge_sequences.npz hand-built MULTI-tic scenarios: a head-on winner that ends a tic alive with health <= 0
                    (game.py:156-165 is an elif chain) at health_dec 9 / 3 / 1, then starves, eats, or wins again
  mcts_<name>.npz   reference Agent + MPGameRunner with a deterministic stub value function: recorded
                    in-tree moves, root moves, root Q, and the full (key, Q, W, N, age) tables per root turn
  funcs.npz         softermax / argmaxs / numpy.random.choice known answers
  pit_<name>.npz    the reference's pit MPGameRunner.run(Alice, Bob, Alice_snake_cnt) with two stand-in value functions:
                    start layouts, spawned food cells, every move, the winner list (incl. the early exit)
  replay_<name>.npz the text frames Game.tic(show=True) appends to replay.rep, with the dumps they were drawn from
"""
import hashlib
import io
import os
import random as pyrandom
import sys
from contextlib import redirect_stdout

import numpy as np

REF = os.environ.get("ASZ_REFERENCE", "/root/reference/code")
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

import utils.game as rg  # noqa: E402
import utils.agent as ra  # noqa: E402
import utils.mp_game_runner as rr  # noqa: E402


# ---------------------------------------------------------------------------------------------
# NumPy >= 2.3 removed ndarray.tostring (agent.py:175): return an ndarray subclass that still has it.
# This does not edit the reference; it only wraps Game.make_state at import time.
class _Sub(np.ndarray):
    def tostring(self):
        return self.tobytes()


_orig_make_state = rg.Game.make_state
rg.Game.make_state = lambda self, you, last_move: _orig_make_state(self, you, last_move).view(_Sub)

# record the food cell picked inside Game.tic (game.py:133)
_spawned = []
_orig_choice = rg.choice


def _rec_choice(seq):
    r = _orig_choice(seq)
    if isinstance(seq, tuple) and len(seq) > 0 and isinstance(seq[0], tuple):
        _spawned.append(r)
    return r


rg.choice = _rec_choice


# ---------------------------------------------------------------------------------------------
def dump(game):
    """canonical dump of a reference Game (same format as oracle og_dump)."""
    H, W, S = game.height, game.width, game.snake_cnt
    snake = np.zeros((S, 6), np.int32)
    owner = np.full(H * W, -1, np.int32)
    dist = np.zeros(H * W, np.int32)
    food = np.zeros(H * W, np.int32)
    for i in range(S):
        snake[i, 3] = game.last_moves[i]
        r = game.rewards[i]
        snake[i, 5] = 0 if r is None else int(r)
        snake[i, 4] = -1
    for s in game.snakes:
        snake[s.id, 0] = 1
        snake[s.id, 1] = s.health
        snake[s.id, 2] = s.length
        hy, hx = s.head.position
        if 0 <= hy < H and 0 <= hx < W:
            snake[s.id, 4] = hy * W + hx
        node, d = s.tail, 1
        while node:
            y, x = node.position
            if 0 <= y < H and 0 <= x < W:
                owner[y * W + x] = s.id
                dist[y * W + x] = d
            node = node.prev_node
            d += 1
    for (y, x) in game.food:
        food[y * W + x] = 1
    counters = np.array([game.wall_collision, game.body_collision, game.head_collision, game.starvation,
                         game.food_eaten, game.game_length, 0, game.id], np.int32)
    return dict(snake=snake, owner=owner, dist=dist, food=food, counters=counters)


def digest(plane):
    return np.frombuffer(hashlib.blake2b(np.ascontiguousarray(plane).tobytes(), digest_size=8).digest(), np.uint64)[0]


def build_ref_game(H, W, S, health_dec, snakes, last_moves, food, chance=0.15, gid=0):
    """snakes: list over ids of None (dead) or (health, [(y,x) head..tail])."""
    g = rg.Game(gid, H, W, S, health_dec, chance)
    g.snakes = []
    for i, spec in enumerate(snakes):
        if spec is None:
            g.rewards[i] = -1.0
            continue
        health, body = spec
        g.snakes.append(rg.Snake(i, health, list(body)))
    g.last_moves = {i: last_moves[i] for i in range(S)}
    g.food = set(food)
    g.heads = {}
    for s in g.snakes:
        g.heads.setdefault(s.head.position, set()).add(s)
    g.bodies = {b for s in g.snakes for b in s}
    g.empty_positions = {(y, x) for y in range(H) for x in range(W)}
    for s in g.snakes:
        g.empty_positions.discard(s.head.position)
        for b in s:
            g.empty_positions.discard(b)
    for f in g.food:
        g.empty_positions.discard(f)
    return g


# ---------------------------------------------------------------------------------------------
def gen_env(name, H, W, S, health_dec, n_games, seed, max_tics=400, full_plane_tics=6):
    pyrandom.seed(seed)
    rng = np.random.default_rng(seed)
    recs = dict(init_start=[], init_last=[], init_food=[], init_nfood=[], game_ptr=[0])
    t_moves, t_spawn, t_nlive, t_ended = [], [], [], []
    d_snake, d_owner, d_dist, d_food, d_cnt = [], [], [], [], []
    digests, dig_ptr = [], [0]
    full_planes, full_idx = [], []
    for gi in range(n_games):
        g = rg.Game(gi, H, W, S, health_dec)
        recs["init_start"].append([s.head.position for s in g.snakes])
        recs["init_last"].append([g.last_moves[i] for i in range(S)])
        fl = sorted(g.food)
        recs["init_nfood"].append(len(fl))
        recs["init_food"].append(fl + [(-1, -1)] * (S + 1 - len(fl)))
        for t in range(max_tics):
            n = len(g.snakes)
            mv = rng.integers(0, 3, size=n).tolist()
            _spawned.clear()
            res = g.tic(list(mv))
            sp = _spawned[-1] if _spawned else None
            t_moves.append(mv + [-1] * (S - n))
            t_nlive.append(n)
            t_spawn.append(-1 if sp is None else sp[0] * W + sp[1])
            t_ended.append(0 if res == 0 else 1)
            d = dump(g)
            d_snake.append(d["snake"]); d_owner.append(d["owner"]); d_dist.append(d["dist"])
            d_food.append(d["food"]); d_cnt.append(d["counters"])
            states = g.get_states() if res == 0 or g.snakes else []
            for k, st in enumerate(states):
                digests.append(digest(st))
                if gi < 3 and t < full_plane_tics:
                    full_planes.append(np.ascontiguousarray(st)); full_idx.append((len(t_moves) - 1, k))
            dig_ptr.append(len(digests))
            if res != 0:
                break
        recs["game_ptr"].append(len(t_moves))
    out = dict(H=H, W=W, S=S, health_dec=health_dec,
               init_start=np.array(recs["init_start"], np.int32), init_last=np.array(recs["init_last"], np.int32),
               init_food=np.array(recs["init_food"], np.int32), init_nfood=np.array(recs["init_nfood"], np.int32),
               game_ptr=np.array(recs["game_ptr"], np.int64),
               moves=np.array(t_moves, np.int8), nlive=np.array(t_nlive, np.int8), spawn=np.array(t_spawn, np.int16),
               ended=np.array(t_ended, np.int8),
               snake=np.array(d_snake, np.int16), owner=np.array(d_owner, np.int8), dist=np.array(d_dist, np.int16),
               food=np.array(d_food, np.int8), counters=np.array(d_cnt, np.int32),
               digests=np.array(digests, np.uint64), dig_ptr=np.array(dig_ptr, np.int64),
               full_planes=np.array(full_planes, np.float32), full_idx=np.array(full_idx, np.int32))
    np.savez_compressed(os.path.join(HERE, "env_%s.npz" % name), **out)
    print("env_%s: %d games, %d tics, %d planes" % (name, n_games, len(t_moves), len(digests)))


# ---------------------------------------------------------------------------------------------
def gen_edge_cases():
    """One reference tic per scenario of SURVEY.md Appendix D (11x11, 4 snake slots)."""
    H = W = 11
    S = 4
    cases = []

    def add(name, snakes, last_moves, food, moves, health_dec=1):
        g = build_ref_game(H, W, S, health_dec, snakes, last_moves, food, chance=0.0)
        before = dump(g)
        planes_before = [np.ascontiguousarray(p) for p in g.get_states()]
        res = g.tic(list(moves))
        after = dump(g)
        planes_after = [np.ascontiguousarray(p) for p in g.get_states()]
        cases.append(dict(name=name, before=before, after=after, moves=list(moves), ended=0 if res == 0 else 1,
                          health_dec=health_dec, planes_before=planes_before, planes_after=planes_after))

    def body(cells):
        return list(cells)

    far = (100, body([(9, 9), (9, 8), (9, 7)]))          # a bystander far away, heading right (last move 1)
    far2 = (100, body([(9, 1), (9, 2), (9, 3)]))         # another bystander heading left (last move 3)
    # D-1 relative move mapping for each last_move value: snake 0 at centre
    for last in range(4):
        for m in range(3):
            tail_dir = {0: (1, 0), 1: (0, -1), 2: (-1, 0), 3: (0, 1)}[last]
            b = [(5, 5), (5 + tail_dir[0], 5 + tail_dir[1]), (5 + 2 * tail_dir[0], 5 + 2 * tail_dir[1])]
            add("relmove_last%d_m%d" % (last, m), [(100, b), far, None, None], [last, 1, 0, 0], [], [m, 1])
    # D-2 moves indexed by live-list position: snake 0 dead, snakes 1 and 3 alive
    add("live_list_index", [None, (100, [(5, 5), (6, 5), (7, 5)]), None, far], [0, 0, 0, 1], [], [0, 2])
    # D-3 head-on on food: equal lengths, lower-listed eats, grows and survives
    add("headon_food_first_come", [(100, [(5, 4), (5, 3), (5, 2)]), (100, [(5, 6), (5, 7), (5, 8)]), far, None],
        [1, 3, 1, 0], [(5, 5)], [1, 1, 1])
    # head-on without food: equal lengths both die
    add("headon_equal_both_die", [(100, [(5, 4), (5, 3), (5, 2)]), (100, [(5, 6), (5, 7), (5, 8)]), far, far2],
        [1, 3, 1, 3], [], [1, 1, 1, 1])
    # head-on, longer wins
    add("headon_longer_wins", [(100, [(5, 4), (5, 3), (5, 2), (5, 1)]), (100, [(5, 6), (5, 7), (5, 8)]), far, None],
        [1, 3, 1, 0], [], [1, 1, 1])
    # three-way head-on, lengths 5,4,3
    add("headon_three_way", [(100, [(5, 4), (5, 3), (5, 2), (5, 1), (5, 0)]), (100, [(5, 6), (5, 7), (5, 8), (5, 9)]),
                             (100, [(4, 5), (3, 5), (2, 5)]), far2], [1, 3, 2, 3], [], [1, 1, 1, 1])
    # D-4 starvation vs eating at health 1
    add("starve_vs_eat", [(1, [(5, 4), (5, 3), (5, 2)]), (1, [(2, 2), (2, 1), (2, 0)]), far, None],
        [1, 1, 1, 0], [(5, 5)], [1, 1, 1])
    add("starve_health_dec9", [(9, [(5, 4), (5, 3), (5, 2)]), (10, [(2, 2), (2, 1), (2, 0)]), far, None],
        [1, 1, 1, 0], [], [1, 1, 1], health_dec=9)
    # D-5 moving into a cell the tail vacates this tic is safe ...
    add("tail_vacated_safe", [(100, [(5, 4), (5, 3), (5, 2)]), (100, [(4, 5), (4, 6), (5, 6), (5, 5)]), far, None],
        [1, 3, 1, 0], [], [1, 1, 1])
    # ... but a stacked tail is a body collision
    add("stacked_tail_body", [(100, [(5, 4), (5, 3), (5, 2)]), (100, [(4, 5), (4, 6), (5, 5), (5, 5)]), far, None],
        [1, 3, 1, 0], [], [1, 1, 1])
    # own-body collision (long snake curling)
    add("self_collision", [(100, [(5, 5), (5, 4), (6, 4), (6, 5), (6, 6), (5, 6), (4, 6)]), far, None, None],
        [1, 1, 0, 0], [], [2, 1])
    # D-6 priority wall > body > head > starvation: starving snake hits the wall -> wall only
    add("priority_wall_over_starve", [(1, [(0, 5), (1, 5), (2, 5)]), far, far2, None], [0, 1, 3, 0], [], [1, 1, 1])
    # body beats head-on: two heads meet on a third snake's body
    add("priority_body_over_head", [(100, [(5, 4), (5, 3), (5, 2)]), (100, [(5, 6), (5, 7), (5, 8)]),
                                    (100, [(3, 5), (4, 5), (5, 5), (6, 5), (7, 5)]), None], [1, 3, 0, 0], [],
        [1, 1, 1])
    # head-on winner with health 1 that does not eat: head-on branch only (elif chain => no starvation check)
    add("headon_winner_low_health", [(1, [(5, 4), (5, 3), (5, 2), (5, 1)]), (100, [(5, 6), (5, 7), (5, 8)]), far, None],
        [1, 3, 1, 0], [], [1, 1, 1])
    # D-8 game start: three stacked segments, first two tics
    add("start_stacked_tic1", [(100, [(1, 1)] * 3), (100, [(9, 9)] * 3), (100, [(9, 1)] * 3), (100, [(1, 9)] * 3)],
        [2, 0, 0, 2], [(5, 5), (2, 2), (8, 8), (8, 2), (2, 8)], [1, 1, 1, 1])
    # eat then grow: tail duplicated, then the duplicate delays the tail by one tic
    add("eat_and_grow", [(50, [(5, 4), (5, 3), (5, 2)]), far, None, None], [1, 1, 0, 0], [(5, 5)], [1, 1])
    add("after_grow_stacked_tail_moves", [(100, [(5, 5), (5, 4), (5, 3), (5, 3)]), far, None, None], [1, 1, 0, 0], [],
        [1, 1])
    # last two snakes die together => game ends with no winner
    add("draw_all_die", [(100, [(5, 4), (5, 3), (5, 2)]), (100, [(5, 6), (5, 7), (5, 8)]), None, None], [1, 3, 0, 0],
        [], [1, 1])
    # one survivor => winner
    add("winner", [(100, [(0, 4), (1, 4), (2, 4)]), (100, [(5, 6), (5, 7), (5, 8)]), None, None], [0, 3, 0, 0], [],
        [1, 1])
    # wall on every side
    for nm, b, last in (("wall_up", [(0, 5), (1, 5), (2, 5)], 0), ("wall_right", [(5, 10), (5, 9), (5, 8)], 1),
                        ("wall_down", [(10, 5), (9, 5), (8, 5)], 2), ("wall_left", [(5, 0), (5, 1), (5, 2)], 3)):
        add(nm, [(100, b), far if nm != "wall_right" else far2, (100, [(2, 2), (2, 3), (2, 4)]), None],
            [last, 1 if nm != "wall_right" else 3, 3, 0], [], [1, 1, 1])
    # dead snake frees its cells; another head may enter a dying snake's body cell -> still body collision
    add("enter_dying_body", [(100, [(0, 5), (1, 5), (2, 5), (3, 5)]), (100, [(2, 4), (2, 3), (2, 2)]), far, None],
        [0, 1, 1, 0], [], [1, 1, 1])
    out = {}
    out["names"] = np.array([c["name"] for c in cases])
    for k in ("snake", "owner", "dist", "food", "counters"):
        out["before_" + k] = np.array([c["before"][k] for c in cases], np.int32)
        out["after_" + k] = np.array([c["after"][k] for c in cases], np.int32)
    out["moves"] = np.array([c["moves"] + [-1] * (S - len(c["moves"])) for c in cases], np.int8)
    out["ended"] = np.array([c["ended"] for c in cases], np.int8)
    out["health_dec"] = np.array([c["health_dec"] for c in cases], np.int32)
    pb, pa, pbp, pap = [], [], [0], [0]
    for c in cases:
        pb += c["planes_before"]; pa += c["planes_after"]
        pbp.append(len(pb)); pap.append(len(pa))
    out["planes_before"] = np.array(pb, np.float32); out["planes_after"] = np.array(pa, np.float32)
    out["pb_ptr"] = np.array(pbp, np.int64); out["pa_ptr"] = np.array(pap, np.int64)
    np.savez_compressed(os.path.join(HERE, "edge_cases.npz"), **out)
    print("edge_cases: %d scenarios" % len(cases))


# ---------------------------------------------------------------------------------------------
def gen_edge_sequences():
    """Multi-tic scenarios (11x11, 4 snake slots, no food spawn).  game.py:156-165 is an elif chain: a snake that shares its
    head cell with a shorter snake takes the head-on branch and skips the starvation check, so it can end a tic alive with
    health <= 0; the next tic it starves (game.py:163-165) unless it eats (health = 100, game.py:123-125) or wins again."""
    H = W = 11
    S = 4
    cases = []

    def add(name, snakes, last_moves, food, move_seq, health_dec):
        g = build_ref_game(H, W, S, health_dec, snakes, last_moves, food, chance=0.0)
        before = dump(g)
        tics = []
        min_health = 1000
        for moves in move_seq:
            n = len(g.snakes)
            assert len(moves) == n, (name, moves, n)
            res = g.tic(list(moves))
            after = dump(g)
            planes = [np.ascontiguousarray(p) for p in g.get_states()] if res == 0 else []   # ended games emit no rows
            for sn in g.snakes:
                min_health = min(min_health, sn.health)
            tics.append(dict(moves=list(moves), after=after, ended=0 if res == 0 else 1, planes=planes))
            if res != 0:
                break
        assert len(tics) == len(move_seq), name
        cases.append(dict(name=name, before=before, tics=tics, health_dec=health_dec, min_health=min_health))

    far = (100, [(9, 6), (9, 5), (9, 4)])          # bystanders that keep the game going for three tics, heading right / left
    far2 = (100, [(7, 4), (7, 5), (7, 6)])
    A4 = [(5, 4), (5, 3), (5, 2), (5, 1)]          # length 4 heading right; meets B3 on (5, 5)
    B3 = [(5, 6), (5, 7), (5, 8)]                  # length 3 heading left
    food_far = (1, 9)                              # a food cell so that the food channel shows (101 - health) * 0.01
    for dec, hp in ((9, 5), (9, 9), (3, 2), (3, 3), (1, 1)):
        tag = "dec%d_hp%d" % (dec, hp)
        # wins the head-on with health hp - dec <= 0, then goes straight without eating: starves on the second tic
        add("neg_health_starves_" + tag, [(hp, A4), (100, B3), far, far2], [1, 3, 1, 3], [food_far],
            [[1, 1, 1, 1], [1, 1, 1], [1, 1]], dec)
        # ... turns onto food on the second tic: back to 100 and grows
        add("neg_health_eats_" + tag, [(hp, A4), (100, B3), far, far2], [1, 3, 1, 3], [food_far, (4, 5)],
            [[1, 1, 1, 1], [0, 1, 1], [1, 1, 1]], dec)
        # ... runs into the wall on the second tic: wall has priority over starvation (one cause, game.py:148-165)
        add("neg_health_wall_" + tag, [(hp, [(0, 4), (0, 3), (0, 2), (0, 1)]), (100, [(0, 6), (0, 7), (0, 8)]), far, far2],
            [1, 3, 1, 3], [food_far], [[1, 1, 1, 1], [0, 1, 1]], dec)
    # two head-on wins in a row: 5 > 3 on (5, 5), then 5 > 4 on (5, 6); health falls twice below zero, then starvation
    A5 = [(5, 4), (5, 3), (5, 2), (5, 1), (5, 0)]
    C4 = [(3, 6), (2, 6), (1, 6), (0, 6)]          # heading down: (4, 6) after one tic, (5, 6) after two
    add("neg_health_twice_dec9", [(5, A5), (100, B3), (100, C4), far2], [1, 3, 2, 3], [food_far],
        [[1, 1, 1, 1], [1, 1, 1], [1, 1]], 9)
    add("neg_health_twice_dec1", [(1, A5), (100, B3), (100, C4), far2], [1, 3, 2, 3], [food_far],
        [[1, 1, 1, 1], [1, 1, 1], [1, 1]], 1)
    # the head-on winner with negative health is the last snake standing: reward +1 with health <= 0
    add("neg_health_wins_game_dec9", [(5, A4), (100, B3), None, None], [1, 3, 0, 0], [food_far], [[1, 1]], 9)
    # health exactly 0 after a head-on win (dec 3, health 3); starves on the next tic
    add("zero_health_then_starves_dec3", [(3, A4), (100, B3), far, None], [1, 3, 1, 0], [], [[1, 1, 1], [0, 1]], 3)
    assert min(c["min_health"] for c in cases) < 0 and any(c["min_health"] == 0 for c in cases)
    out = dict(names=np.array([c["name"] for c in cases]), health_dec=np.array([c["health_dec"] for c in cases], np.int32),
               min_health=np.array([c["min_health"] for c in cases], np.int32))
    for k in ("snake", "owner", "dist", "food", "counters"):
        out["before_" + k] = np.array([c["before"][k] for c in cases], np.int32)
    tic_ptr, mv, ended, planes, pl_ptr = [0], [], [], [], [0]
    after = {k: [] for k in ("snake", "owner", "dist", "food", "counters")}
    for c in cases:
        for t in c["tics"]:
            mv.append(t["moves"] + [-1] * (S - len(t["moves"])))
            ended.append(t["ended"])
            for k in after:
                after[k].append(t["after"][k])
            planes += t["planes"]
            pl_ptr.append(len(planes))
        tic_ptr.append(len(mv))
    out.update(tic_ptr=np.array(tic_ptr, np.int64), moves=np.array(mv, np.int8), ended=np.array(ended, np.int8),
               planes=np.array(planes, np.float32), pl_ptr=np.array(pl_ptr, np.int64))
    for k in after:
        out["after_" + k] = np.array(after[k], np.int32)
    np.savez_compressed(os.path.join(HERE, "edge_sequences.npz"), **out)
    print("edge_sequences: %d scenarios, %d tics, min health %d" % (len(cases), len(mv), min(c["min_health"] for c in cases)))


# ---------------------------------------------------------------------------------------------
# Deterministic stub value function, defined on the plane bytes; the oracle (og_plane_key, og_stub_value,
# og_obstacle_mask) and the CUDA engine implement the same definition independently.
M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _fmix64(k):
    k = k ^ (k >> np.uint64(33))
    k = k * np.uint64(0xff51afd7ed558ccd)
    k = k ^ (k >> np.uint64(33))
    k = k * np.uint64(0xc4ceb9fe1a85ec53)
    k = k ^ (k >> np.uint64(33))
    return k


def plane_keys(X):
    """X: (n, h, w, 3) float32 -> (n, 2) uint64."""
    X = np.ascontiguousarray(X, dtype=np.float32)
    n = X.shape[0]
    u = X.view(np.uint32).reshape(n, -1, 3).astype(np.uint64)
    a, b, c = u[..., 0], u[..., 1], u[..., 2]
    p = np.arange(u.shape[1], dtype=np.uint64)[None, :]
    wall = (a == 0) & (b == np.uint64(0x3F800000)) & (c == 0)
    x = (a << np.uint64(32)) | b
    y = (c << np.uint64(32)) | p
    with np.errstate(over="ignore"):
        h0 = _fmix64(_fmix64(y ^ np.uint64(0x9E3779B97F4A7C15)) ^ x)
        h1 = _fmix64(_fmix64(x ^ np.uint64(0xC2B2AE3D27D4EB4F)) + y)
        h0 = np.where(wall, np.uint64(0), h0).sum(axis=1, dtype=np.uint64)
        h1 = np.where(wall, np.uint64(0), h1).sum(axis=1, dtype=np.uint64)
    h0 = np.where(h0 == 0, np.uint64(1), h0)
    return np.stack([h0, h1], axis=1)


class StubNet:
    """AlphaNNet.v contract (alpha_nnet.py:61-76): value from the plane key, then the obstacle mask."""

    def __init__(self):
        self.calls = []

    def v(self, X):
        X = np.array(X)
        keys = plane_keys(X)
        V = np.zeros((len(X), 3), np.float32)
        for i in range(3):
            xs = ((keys[:, 1] >> np.uint64(16 * i)) & np.uint64(0xFFFF)).astype(np.float32)
            V[:, i] = (xs - np.float32(32767.5)) * np.float32(1.0 / 32768.0)
        cy, cx = X.shape[1] // 2, X.shape[2] // 2
        thr = np.float32(0.04)
        V[X[:, cy, cx - 1, 1] >= thr, 0] = -1.0
        V[X[:, cy - 1, cx, 1] >= thr, 1] = -1.0
        V[X[:, cy, cx + 1, 1] >= thr, 2] = -1.0
        self.calls.append(len(X))
        return V


def gen_mcts(name, H, W, S, health_dec, G, base, training, D, breadth, root_turns, seed, custom=None, warm_tics=0,
             need_nonpositive_health=False):
    """custom: list of G (snakes, last_moves, food) specs for build_ref_game -- the run starts from hand-built states (the
    tests then take the start state from the t0_before_* dumps: `custom` = 1 in the fixture).
    warm_tics: uniform-random root tics played before the search starts (mid-game start: fewer live snakes, deeper
    sub-games at 19x19x8, agent.py:45); games that end while warming up are re-created."""
    pyrandom.seed(seed)
    np.random.seed(seed)
    net = StubNet()
    agent = ra.Agent(net, base, training, D, breadth)
    parallel = min(8, breadth)
    epochs = breadth // parallel
    if custom is not None:
        assert len(custom) == G
        games = {i: build_ref_game(H, W, S, health_dec, custom[i][0], custom[i][1], custom[i][2], chance=0.15, gid=i) for i in range(G)}
    else:
        games = {i: rg.Game(i, H, W, S, health_dec) for i in range(G)}
    if warm_tics:
        wr = np.random.default_rng(seed + 1000)
        for i in range(G):
            while True:
                g = rg.Game(i, H, W, S, health_dec)
                ok = True
                for _ in range(warm_tics[i] if isinstance(warm_tics, (list, tuple)) else warm_tics):
                    if g.tic(wr.integers(0, 3, size=len(g.snakes)).tolist()) != 0:
                        ok = False
                        break
                if ok:
                    break
            games[i] = g
            # counters restart: the fixtures compare the six log counters from the start of the recorded run
            g.wall_collision = g.body_collision = g.head_collision = g.starvation = g.food_eaten = g.game_length = 0
    nonpos = dict(n=0)
    orig_tic = rg.Game.tic

    def tic_probe(self, moves, show=False):
        r = orig_tic(self, moves, show)
        nonpos["n"] += sum(1 for sn in self.snakes if sn.health <= 0)
        return r
    rg.Game.tic = tic_probe
    if custom is not None or warm_tics:
        init = dict(start=np.zeros((G, S, 2)), last=np.zeros((G, S)), food=np.zeros((G, S + 1, 2)), nfood=np.zeros(G))
    else:
        init = dict(start=[[s.head.position for s in games[i].snakes] for i in range(G)],
                    last=[[games[i].last_moves[k] for k in range(S)] for i in range(G)],
                    food=[sorted(games[i].food) + [(-1, -1)] * (S + 1 - len(games[i].food)) for i in range(G)],
                    nfood=[len(games[i].food) for i in range(G)])

    # hooks: capture (ids, moves) of every MCTSAgent.make_moves call, epoch boundaries via MCTSMPGameRunner.run
    calls = []
    orig_mm = ra.MCTSAgent.make_moves
    orig_run = rr.MCTSMPGameRunner.run
    state = dict(epoch=-1, step=0)

    def mm(self, subgames, ids):
        mv = orig_mm(self, subgames, ids)
        calls.append((state["epoch"], state["step"], list(ids), [int(m) for m in mv]))
        state["step"] += 1
        return mv

    def run(self, alice, depth):
        state["epoch"] += 1
        state["step"] = 0
        return orig_run(self, alice, depth)

    ra.MCTSAgent.make_moves = mm
    rr.MCTSMPGameRunner.run = run
    per_turn = []
    try:
        for turn in range(root_turns):
            if not games:
                break
            ids = []
            for gid in games:
                ids += games[gid].get_ids()
            before = {gid: dump(games[gid]) for gid in games}
            calls.clear(); state["epoch"] = -1
            with redirect_stdout(io.StringIO()):
                moves = agent.make_moves(games, ids)
            # trace of in-tree moves with absolute sub-game ids: the reference numbers sub-games consecutively over
            # the LIVE games (agent.py:42-50); translate to game_id*parallel + sibling
            live = list(games.keys())
            tree = np.full((epochs, max(D, 1), G * parallel, S), 255, np.uint8)
            for (ep, st, cids, cmv) in calls:
                for (sub, snake), m in zip(cids, cmv):
                    abs_sub = live[sub // parallel] * parallel + sub % parallel
                    tree[ep, st, abs_sub, snake] = m
            # root Q as read by Agent.make_moves (agent.py:83-87): look the root keys up in the table now
            q = []
            for gid in games:
                for stt in games[gid].get_states():
                    q.append(np.array(agent.cached_values[stt.tostring()], np.float32).copy())
            # table snapshot BEFORE eviction is not observable; this is after eviction (agent.py:101-110)
            keys_bytes = list(agent.cached_values.keys())
            planes = np.frombuffer(b"".join(keys_bytes), np.float32).reshape(len(keys_bytes), 2 * H - 1, 2 * W - 1, 3)
            tk = plane_keys(planes)
            tQ = np.array([agent.cached_values[k] for k in keys_bytes], np.float32)
            tW = np.array([agent.total_rewards[k] for k in keys_bytes], np.float32)
            tN = np.array([agent.visit_cnts[k] for k in keys_bytes], np.float32)
            tA = np.array([agent.cache_hit[k] for k in keys_bytes], np.int32)
            order = np.lexsort((tk[:, 1], tk[:, 0]))
            per_turn.append(dict(ids=np.array(ids, np.int32), root_moves=np.array(moves, np.uint8), root_q=np.array(q, np.float32),
                                 tree=tree, tab_keys=tk[order], tab_Q=tQ[order], tab_W=tW[order], tab_N=tN[order],
                                 tab_age=tA[order], live=np.array(live, np.int32),
                                 before={k: np.array([before[g][k] for g in live]) for k in ("snake", "owner", "dist", "food", "counters")},
                                 n_evals=np.array(sum(net.calls))))
            # root tic (mp_game_runner.py:44-66)
            mfg = {gid: [] for gid in games}
            for i, m in enumerate(moves):
                mfg[ids[i][0]].append(m)
            spawn = {}
            for gid in list(games.keys()):
                _spawned.clear()
                res = games[gid].tic(mfg[gid])
                spawn[gid] = -1 if not _spawned else _spawned[-1][0] * W + _spawned[-1][1]
                if res != 0:
                    del games[gid]
            per_turn[-1]["spawn"] = np.array([spawn[g] for g in live], np.int32)
    finally:
        ra.MCTSAgent.make_moves = orig_mm
        rr.MCTSMPGameRunner.run = orig_run
        rg.Game.tic = orig_tic
    if need_nonpositive_health:
        assert nonpos["n"] > 0, "no snake ended a (sub-)game tic alive with health <= 0"
    out = dict(H=H, W=W, S=S, health_dec=health_dec, G=G, base=base, training=int(training), D=D, breadth=breadth,
               custom=int(custom is not None or bool(warm_tics)), nonpositive_health_tics=nonpos["n"],
               n_turns=len(per_turn), init_start=np.array(init["start"], np.int32), init_last=np.array(init["last"], np.int32),
               init_food=np.array(init["food"], np.int32), init_nfood=np.array(init["nfood"], np.int32))
    for t, d in enumerate(per_turn):
        for k, v in d.items():
            if k == "before":
                for kk, vv in v.items():
                    out["t%d_before_%s" % (t, kk)] = vv.astype(np.int32)
            else:
                out["t%d_%s" % (t, k)] = v
    if training:
        out["records"] = np.array([np.ascontiguousarray(r) for r in agent.records], np.float32)
    np.savez_compressed(os.path.join(HERE, "mcts_%s.npz" % name), **out)
    print("mcts_%s: %d root turns, %d evals, table %d, %d snake-tics with health <= 0" % (
        name, len(per_turn), sum(net.calls), len(agent.cached_values), nonpos["n"]))


def gen_funcs():
    rng = np.random.default_rng(7)
    Z = rng.uniform(-0.999, 0.999, size=(256, 3)).astype(np.float32)
    Z[:40][rng.random((40, 3)) < 0.4] = -1.0
    Z[0] = [-1, -1, -1]
    Z[1] = [0.5, 0.5, 0.5]; Z[2] = [0.5, 0.5, 0.2]; Z[3] = [0.2, 0.5, 0.5]; Z[4] = [0.5, 0.2, 0.5]
    sm = {}
    for base in (2, 3, 10, 100):
        a = ra.Agent(None, base)
        sm[base] = np.array([np.asarray(a.softermax(z), np.float32) for z in Z], np.float32)
    am = np.array(ra.Agent(None).argmaxs(list(Z)), np.int32)
    # numpy.random.choice: index as a function of (p, u) -- replace the uniform draw to learn u
    us = rng.random(256)
    idx = []
    for z, u in zip(sm[2], us):
        cdf = np.cumsum(z.astype(np.float64)); cdf /= cdf[-1]
        idx.append(int(cdf.searchsorted(u, side="right")))
    # and check that restatement against the real choice() using a generator whose draw we know
    np.random.seed(123)
    st0 = np.random.get_state()
    real, pred = [], []
    for z in sm[2][:64]:
        np.random.set_state(st0); u = np.random.random_sample()
        np.random.set_state(st0); real.append(int(ra.choice([0, 1, 2], p=z)))
        cdf = np.cumsum(z.astype(np.float64)); cdf /= cdf[-1]
        pred.append(int(cdf.searchsorted(u, side="right")))
        st0 = np.random.get_state()
    assert real == pred, "choice() restatement does not match numpy.random.choice"
    np.savez_compressed(os.path.join(HERE, "funcs.npz"), Z=Z, softermax_2=sm[2], softermax_3=sm[3], softermax_10=sm[10],
                        softermax_100=sm[100], argmaxs=am, choice_u=us, choice_idx=np.array(idx, np.int32))
    print("funcs: ok")


def gen_replay(name, H, W, S, health_dec, n_games, seed, max_tics=400):
    """Game.tic(moves, show=True) (game.py:140-141,194-195,281-300): the two text frames per tic the reference appends to
    replay.rep, next to the pre-tic dump, the moves and the post-tic dump they were drawn from."""
    import tempfile
    pyrandom.seed(seed)
    rng = np.random.default_rng(seed)
    keys = ("snake", "owner", "dist", "food")
    pre = {k: [] for k in keys}; post = {k: [] for k in keys}
    t_moves, texts = [], []
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            for gi in range(n_games):
                g = rg.Game(gi, H, W, S, health_dec)
                for t in range(max_tics):
                    n = len(g.snakes)
                    mv = rng.integers(0, 3, size=n).tolist()
                    d0 = dump(g)
                    if os.path.exists("replay.rep"):
                        os.remove("replay.rep")
                    res = g.tic(list(mv), True)
                    d1 = dump(g)
                    texts.append(open("replay.rep").read())
                    for k in keys:
                        pre[k].append(d0[k]); post[k].append(d1[k])
                    t_moves.append(mv + [-1] * (S - n))
                    if res != 0:
                        break
        finally:
            os.chdir(cwd)
    blob = "\x00".join(texts).encode()
    out = dict(H=H, W=W, S=S, moves=np.array(t_moves, np.int8), text=np.frombuffer(blob, np.uint8))
    for k in keys:
        out["pre_" + k] = np.array(pre[k], np.int16); out["post_" + k] = np.array(post[k], np.int16)
    np.savez_compressed(os.path.join(HERE, "replay_%s.npz" % name), **out)
    print("replay_%s: %d games, %d tics" % (name, n_games, len(t_moves)))


def gen_pit(name, H, S, health_dec, G, alice_cnt, seed):
    """pit_mp_game_runner.py:14-63 with pit_agent.py:10-13 agents over two different stand-in value functions"""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from tests.helpers import KeyStubNet
    import utils.pit_agent as rpa
    import utils.pit_mp_game_runner as rp
    pyrandom.seed(seed)
    np.random.seed(seed)
    runner = rp.MPGameRunner(H, H, S, health_dec, G)
    games = runner.games
    init = dict(start=[[s.head.position for s in games[i].snakes] for i in range(G)],
                last=[[games[i].last_moves[k] for k in range(S)] for i in range(G)],
                food=[sorted(games[i].food) + [(-1, -1)] * (S + 1 - len(games[i].food)) for i in range(G)],
                nfood=[len(games[i].food) for i in range(G)])
    spawn, moves = {}, {}
    orig_tic = rg.Game.tic

    def tic(self, mv, show=False):
        t = self.game_length
        live = [sn.id for sn in self.snakes]
        _spawned.clear()
        r = orig_tic(self, mv, show)
        spawn[(self.id, t)] = -1 if not _spawned else _spawned[-1][0] * H + _spawned[-1][1]
        for sid, m in zip(live, mv):
            moves[(self.id, t, sid)] = int(m)
        return r
    rg.Game.tic = tic
    try:
        Alice, Bob = rpa.Agent(KeyStubNet(1)), rpa.Agent(KeyStubNet(0))
        winners = runner.run(Alice, Bob, alice_cnt)
    finally:
        rg.Game.tic = orig_tic
    T = max(t for (_, t) in spawn) + 1
    sp = np.full((T, G), -2, np.int32)          # -2: the game was no longer running that turn
    mv = np.full((T, G, 8), 255, np.uint8)
    for (g, t), c in spawn.items():
        sp[t, g] = c
    for (g, t, sid), m in moves.items():
        mv[t, g, sid] = m
    out = dict(H=H, S=S, health_dec=health_dec, G=G, alice_cnt=-1 if alice_cnt is None else alice_cnt,
               winners=np.array([-1 if w is None else w for w in winners], np.int32), spawn=sp, moves=mv,
               init_start=np.array(init["start"], np.int32), init_last=np.array(init["last"], np.int32),
               init_food=np.array(init["food"], np.int32), init_nfood=np.array(init["nfood"], np.int32))
    np.savez_compressed(os.path.join(HERE, "pit_%s.npz" % name), **out)
    early = sum(1 for g in range(G) if winners[g] is not None and (sp[:, g] != -2).sum() > 0)
    print("pit_%s: %d games, %d turns, winners %s" % (name, G, T, np.bincount(out["winners"] + 1, minlength=S + 1).tolist()))


def neg_health_games():
    """four root games in which snake 0 (length 4, low health) faces snake 1 (length 3) two cells away: in every sub-game
    where both go straight, snake 0 wins the head-on and lives on with health <= 0 (game.py:156-165)"""
    far = (100, [(9, 6), (9, 5), (9, 4)])
    far2 = (100, [(7, 4), (7, 5), (7, 6)])
    A4 = [(5, 4), (5, 3), (5, 2), (5, 1)]
    B3 = [(5, 6), (5, 7), (5, 8)]
    A4v = [(4, 5), (3, 5), (2, 5), (1, 5)]         # the same meeting, vertical: heading down / up
    B3v = [(6, 5), (7, 5), (8, 5)]
    return [([(5, A4), (100, B3), far, far2], [1, 3, 1, 3], [(1, 9), (4, 5)]),
            ([(9, A4), (100, B3), far, None], [1, 3, 1, 0], [(1, 9)]),
            ([(2, A4v), (100, B3v), None, (100, [(9, 9), (9, 8), (9, 7)])], [2, 0, 0, 1], [(1, 1), (5, 6)]),
            ([(100, B3), (7, A4), far, far2], [3, 1, 1, 3], [(1, 9), (6, 5)])]


def gen_round2():
    """fixtures added in round 2 (all from the unmodified reference)"""
    gen_edge_sequences()
    # sub-games in which a head-on winner lives on with health <= 0, at the trainer's health_dec 9 and 3
    gen_mcts("11x11x4_neghealth_dec9", 11, 11, 4, 9, G=4, base=2, training=True, D=8, breadth=24, root_turns=4, seed=31,
             custom=neg_health_games(), need_nonpositive_health=True)
    gen_mcts("11x11x4_neghealth_dec3", 11, 11, 4, 3, G=4, base=3, training=True, D=8, breadth=16, root_turns=3, seed=32,
             custom=neg_health_games(), need_nonpositive_health=True)
    # 19x19x8 from a mid-game start: <= 5 live snakes, so sub-games are deeper than one tic (agent.py:45), with deaths,
    # evictions (agent.py:101-110) and > 1,000 evaluations
    gen_pit("1v1", 11, 2, 1, G=40, alice_cnt=None, seed=41)
    gen_pit("2v2", 11, 4, 1, G=40, alice_cnt=None, seed=42)
    gen_pit("1v3", 11, 4, 1, G=30, alice_cnt=1, seed=43)
    gen_pit("3v1_7x7", 7, 4, 3, G=30, alice_cnt=3, seed=44)
    gen_mcts("19x19x8_mid", 19, 19, 8, 1, G=3, base=2, training=True, D=8, breadth=16, root_turns=12, seed=33,
             warm_tics=[10, 18, 30])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "round2":
        gen_round2()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "replay":      # only the fixtures added after the first generation
        gen_replay("11x11x4", 11, 11, 4, 1, 25, seed=21)
        gen_replay("7x7x8", 7, 7, 8, 1, 10, seed=22)
        sys.exit(0)
    gen_funcs()
    gen_edge_cases()
    gen_env("11x11x4", 11, 11, 4, 1, 200, seed=1)
    gen_env("11x11x4_dec9", 11, 11, 4, 9, 60, seed=2)
    gen_env("7x7x4", 7, 7, 4, 1, 60, seed=3)
    gen_env("19x19x8", 19, 19, 8, 1, 30, seed=4)
    gen_env("11x11x2", 11, 11, 2, 3, 40, seed=5)
    gen_env("7x7x8", 7, 7, 8, 1, 40, seed=6)
    gen_mcts("11x11x4_train", 11, 11, 4, 1, G=3, base=2, training=True, D=8, breadth=16, root_turns=12, seed=11)
    gen_mcts("11x11x4_eval", 11, 11, 4, 1, G=2, base=100, training=False, D=8, breadth=24, root_turns=14, seed=12)
    gen_mcts("7x7x4_dec9", 7, 7, 4, 9, G=2, base=3, training=True, D=4, breadth=8, root_turns=10, seed=13)
    gen_mcts("19x19x8", 19, 19, 8, 1, G=1, base=2, training=True, D=8, breadth=8, root_turns=3, seed=14)
    gen_mcts("11x11x4_b5", 11, 11, 4, 1, G=2, base=2, training=True, D=8, breadth=5, root_turns=4, seed=15)
    gen_round2()
    gen_replay("11x11x4", 11, 11, 4, 1, 25, seed=21)
    gen_replay("7x7x8", 7, 7, 8, 1, 10, seed=22)
