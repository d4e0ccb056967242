"""asz_env_step_host (the e2e path of bench.py) and the device-resident path side by side, 300-step blocks.
ASZ_ENV_HINTS_HOST=0|1 selects the L2 policies of the host-buffer path."""
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
kw = dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True)
def step(i):
    _lib.check(L.asz_env_step_host(eng.h, flags, _lib.SPAWN_NATIVE, C.c_void_p(pool[i % 8].data_ptr()), None, C.c_void_p(h_ended.data_ptr()),
                                   C.c_void_p(h_rewards.data_ptr()), C.byref(rows), None, None, eng.stream))
def t(fn, n=300):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
for i in range(20): step(i)
print("host hints=%s |" % os.environ.get("ASZ_ENV_HINTS_HOST", "default"), "e2e", " ".join("%.1f" % t(step) for _ in range(3)),
      "| device", " ".join("%.1f" % t(lambda i: eng.step(**kw)) for _ in range(2)), "| e2e", " ".join("%.1f" % t(step) for _ in range(3)))
