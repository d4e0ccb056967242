"""Fast vs slow state of env_step_kernel under ncu (single-pass metrics, no cache control): device path (fast), then e2e steps
(which write the engine-owned plane buffer), then the device path again (slow)."""
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
kw = dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True)
for i in range(30): eng.step(**kw)       # launches 0..29: device path, fast state
torch.cuda.synchronize()
for i in range(30): step(i)              # launches 30..59: e2e path
for i in range(30): eng.step(**kw)       # launches 60..89: device path, slow state
torch.cuda.synchronize()
