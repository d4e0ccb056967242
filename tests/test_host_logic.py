"""CPU: host-side logic of the mirror that needs no GPU."""
import numpy as np


def test_learning_rate_schedule_matches_keras_piecewise_constant_decay():
    """alpha_nnet.py:79-84: PiecewiseConstantDecay([20, 40, 60, 80, 100], [lr, lr/4, lr/16, lr/64, lr/256, 0]).  Keras returns
    values[0] for step <= 20, values[i] for boundaries[i-1] < step <= boundaries[i], the last value beyond; the optimizer's
    step counter is 0 for the first update."""
    from alphasnake_zero_b200.training import lr_at
    lr = 1e-4
    boundaries = [20, 40, 60, 80, 100]
    values = [lr * 0.25 ** i for i in range(5)] + [0.0]

    def keras(step):
        for b, v in zip(boundaries, values):
            if step <= b:
                return v
        return values[-1]
    for step in range(0, 140):
        assert lr_at(step, lr) == keras(step), step
    assert lr_at(20, lr) == lr and lr_at(21, lr) == lr / 4 and lr_at(100, lr) == lr / 256 and lr_at(101, lr) == 0.0


def test_shard_ranges_cover_all_games():
    from alphasnake_zero_b200.parallel import shard_range
    for total, world in ((32768, 8), (10, 3), (7, 8)):
        got = [shard_range(total, r, world) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
        sizes = [hi - lo for lo, hi in got]
        assert max(sizes) - min(sizes) <= 1


def test_agent_host_helpers_match_reference_known_answers():
    """Agent.softermax / argmaxs host restatements (agent.py:114-137) against the reference's known answers"""
    from alphasnake_zero_b200.utils.agent import Agent
    from tests.helpers import load
    z = load("funcs.npz")
    for base in (2, 3, 10, 100):
        a = Agent(None, base)
        got = np.array([a.softermax(zz) for zz in z["Z"]], np.float32)
        np.testing.assert_allclose(got, z["softermax_%d" % base], rtol=1e-6, atol=1e-7)
    assert Agent(None).argmaxs(list(z["Z"])) == z["argmaxs"].tolist()
