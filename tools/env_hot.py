"""Scan of the hot scheduling word's candidate addresses with the engine's controller off (ASZ_AUTO_CONDITION=0): for each
candidate k, the time per tic + encode launch (65,536 games) right after an L2 read sweep and right after a 1 GB fill.
  [ASZ_LIB=tools/libasz_b200_<variant>.so] python tools/env_hot.py [flags] [n_candidates]      flags: big1 = 256 MB allocated first"""
import os
import sys

os.environ["ASZ_AUTO_CONDITION"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from alphasnake_zero_b200 import _lib  # noqa: E402
from alphasnake_zero_b200.engine import Engine  # noqa: E402

flags = sys.argv[1].split(",") if len(sys.argv) > 1 and sys.argv[1] not in ("", "-") else []
n_k = int(sys.argv[2]) if len(sys.argv) > 2 else 12
torch.cuda.set_device(0)
for f in flags:
    if f.startswith("big"):
        keep = torch.empty(int(f[3:] or 1) << 26, dtype=torch.float32, device="cuda")
eng = Engine(side=11, snakes=4, health_dec=1, games=65536, seed=1)
eng.reset()
_ = eng.planes
junk = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
kw = dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True)


def t(n=80):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.step(**kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000


for _ in range(30):
    eng.step(**kw)
out = []
for k in range(n_k):
    _lib.check(_lib.lib().asz_internal_set_hot_word(eng.h, k))
    eng.condition_l2()
    a = t()
    a2 = t()
    junk.fill_(1.0)
    b = t()
    b2 = t()
    out.append("k%d %.0f/%.0f|%.0f/%.0f" % (k, a, a2, b, b2))
print("%s %s: swept/swept | filled/filled us per launch:  %s" % (os.environ.get("ASZ_LIB", "default"), ",".join(flags), "  ".join(out)))
