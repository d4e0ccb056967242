"""e2e leg of bench.py in isolation: asz_env_step_host with pinned host buffers, per-300-step timings."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alphasnake_zero_b200 import _lib
from alphasnake_zero_b200.engine import Engine
G = 65536
eng = Engine(side=11, snakes=4, health_dec=1, games=G, seed=1); eng.reset(); _ = eng.planes
rng = np.random.default_rng(0)
pool = [torch.from_numpy(rng.integers(0, 3, size=(G, 8), dtype=np.uint8)).pin_memory() for _ in range(8)]
h_ended = torch.zeros(G, dtype=torch.uint8).pin_memory(); h_rewards = torch.zeros(G, 8, dtype=torch.int8).pin_memory()
rows = C.c_int32(0); L = _lib.lib(); flags = _lib.STEP_TIC | _lib.STEP_ENCODE | _lib.STEP_AUTO_RESET
def step(i):
    _lib.check(L.asz_env_step_host(eng.h, flags, _lib.SPAWN_NATIVE, C.c_void_p(pool[i % 8].data_ptr()), None, C.c_void_p(h_ended.data_ptr()),
                                   C.c_void_p(h_rewards.data_ptr()), C.byref(rows), None, None, eng.stream))
def t(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): step(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
kw = dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True)
def td(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): eng.step(**kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
for i in range(20): step(i)
print("e2e   ", " ".join("%.1f" % t(300) for _ in range(4)), "rows", rows.value)
print("device", " ".join("%.1f" % td(300) for _ in range(3)))
print("e2e   ", " ".join("%.1f" % t(300) for _ in range(4)), "rows", rows.value)
act = pool[0].cuda()
def tdev(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): eng.step(actions=act, spawn_mode=2, tic=True, encode=True, auto_reset=True)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
print("device, given actions", " ".join("%.1f" % tdev(300) for _ in range(3)))
import time
print("--- kicks while slow")
x = torch.empty(1 << 29, dtype=torch.float32, device="cuda")
def rep(name):
    print("%-44s device %s" % (name, " ".join("%.1f" % td(200) for _ in range(3))))
rep("now")
time.sleep(0.5); rep("after 0.5 s sleep")
eng.planes.zero_(); torch.cuda.synchronize(); rep("after zeroing the planes buffer")
for _ in range(20): x.fill_(0.0)
torch.cuda.synchronize(); rep("after 20 x fill_ of 2 GB")
eng2 = Engine(side=11, snakes=4, health_dec=1, games=G, seed=2); eng2.reset(); _ = eng2.planes
for _ in range(100): eng2.step(**kw)
torch.cuda.synchronize(); rep("after 100 steps of a second engine")
print("--- triggers while fast")
for i in range(50):
    eng.step(**kw); torch.cuda.synchronize()
rep("after 50 device steps with a sync each")
for i in range(50): step(i)
rep("after 50 e2e steps")
