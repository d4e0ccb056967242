"""Shared helpers for the parity tests."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def digest(plane):
    return np.frombuffer(hashlib.blake2b(np.ascontiguousarray(plane).tobytes(), digest_size=8).digest(), np.uint64)[0]


def assert_dump_equal(got, want, ctx=""):
    """canonical dumps: last_move is only meaningful for live snakes; episode/game_id are engine bookkeeping."""
    gs, ws = np.asarray(got["snake"]).astype(np.int64), np.asarray(want["snake"]).astype(np.int64)
    alive = ws[:, 0] == 1
    assert np.array_equal(gs[:, 0], ws[:, 0]), "alive %s\n%s\n%s" % (ctx, gs, ws)
    assert np.array_equal(gs[alive][:, 1:5], ws[alive][:, 1:5]), "snake table %s\n%s\n%s" % (ctx, gs, ws)
    assert np.array_equal(gs[:, 5], ws[:, 5]), "rewards %s\n%s\n%s" % (ctx, gs, ws)
    for k in ("owner", "dist", "food"):
        assert np.array_equal(np.asarray(got[k]).astype(np.int64), np.asarray(want[k]).astype(np.int64)), "%s %s" % (k, ctx)
    assert np.array_equal(np.asarray(got["counters"])[:6], np.asarray(want["counters"])[:6]), "counters %s %s %s" % (
        ctx, got["counters"], want["counters"])


def golden_dump(z, t):
    return dict(snake=z["snake"][t], owner=z["owner"][t], dist=z["dist"][t], food=z["food"][t], counters=z["counters"][t])
