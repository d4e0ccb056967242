import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphasnake_zero_b200 import _lib
from alphasnake_zero_b200.engine import Engine
eng = Engine(side=11, snakes=4, health_dec=1, games=65536, seed=1)
eng.reset(); _ = eng.planes
def t(kw, n=100):
    for _ in range(10): eng.step(**kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): eng.step(**kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
for _ in range(40): eng.step(spawn_mode=2, tic=True, encode=False, auto_reset=True, random_actions=True)
for rep in range(3):
    print("tic + encode  %.1f us, rows %d" % (t(dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True), 300), int(eng.row_count.item())))
print("tic only      %.1f us" % t(dict(spawn_mode=2, tic=True, encode=False, auto_reset=True, random_actions=True)))
print("encode only   %.1f us" % t(dict(tic=False, encode=True)))
print("tic + encode  %.1f us" % t(dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True)))
print("tic + encode + keys %.1f us" % t(dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True, keys=True)))
x = torch.empty(993462041 // 4, dtype=torch.float32, device="cuda")
def fill():
    x.fill_(1.0)
for _ in range(3): fill()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): fill()
e1.record(); torch.cuda.synchronize()
print("torch fill_ of 993 MB %.1f us = %.0f GB/s" % (e0.elapsed_time(e1) / 20 * 1000, 993.46 / (e0.elapsed_time(e1) / 20)))
