import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphasnake_zero_b200 import _lib
from alphasnake_zero_b200.engine import Engine
G = 65536
eng = Engine(side=11, snakes=4, health_dec=1, games=G, seed=1); eng.reset(); _ = eng.planes
kw = dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True)
kw_enc = dict(tic=False, encode=True)
def probe(name, kw, n=200):
    for _ in range(20): eng.step(**kw)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    t0 = time.perf_counter()
    ev[0].record()
    for i in range(n):
        eng.step(**kw); ev[i + 1].record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    d = np.array([ev[i].elapsed_time(ev[i + 1]) * 1000 for i in range(n)])
    print("%-28s cpu enqueue %.1f us/step, wall %.1f us/step | gpu per step: median %.1f min %.1f p90 %.1f max %.1f" %
          (name, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6, np.median(d), d.min(), np.percentile(d, 90), d.max()))
probe("A tic+encode", kw)
probe("B encode only", kw_enc)
probe("C tic+encode", kw)
probe("C tic+encode", kw)
eng.planes.zero_()
probe("D tic+encode after zero_", kw)
