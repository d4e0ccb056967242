import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphasnake_zero_b200.engine import Engine
G = 65536
eng = Engine(side=11, snakes=4, health_dec=1, games=G, seed=1); eng.reset(); _ = eng.planes
kw = dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True)
kw_enc = dict(tic=False, encode=True)
def t(kw, n=300, pre=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        if pre is not None: pre()
        eng.step(**kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
for _ in range(20): eng.step(**kw)
print("A", " ".join("%.1f" % t(kw) for _ in range(2)))
print("B", " ".join("%.1f" % t(kw_enc) for _ in range(1)))
print("C", " ".join("%.1f" % t(kw) for _ in range(2)))
for mb in (16, 64, 256):
    x = torch.empty(mb * (1 << 18), dtype=torch.float32, device="cuda")
    print("C with a %d MB fill before every launch:" % mb, " ".join("%.1f" % t(kw, 300, lambda: x.fill_(0.0)) for _ in range(2)),
          " then plain C:", " ".join("%.1f" % t(kw) for _ in range(2)))
    print("B", "%.1f" % t(kw_enc), " C", " ".join("%.1f" % t(kw) for _ in range(2)))
