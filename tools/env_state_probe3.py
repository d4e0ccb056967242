"""Phase timings of env_step_kernel from a known-good start: reset (streaming write), A = device tic+encode, E = e2e steps,
B = encode only."""
import ctypes as C, os, sys, time
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
def e2e(i):
    _lib.check(L.asz_env_step_host(eng.h, flags, _lib.SPAWN_NATIVE, C.c_void_p(pool[i % 8].data_ptr()), None, C.c_void_p(h_ended.data_ptr()),
                                   C.c_void_p(h_rewards.data_ptr()), C.byref(rows), None, None, eng.stream))
kw = dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True)
kw_enc = dict(tic=False, encode=True)
import subprocess
def smi():
    q = "clocks.sm,clocks.gr,clocks.mem,clocks.video,power.draw,pstate"
    return subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=" + q, "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
def t(fn, n=600):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n // 2): fn(i)
    s = smi()
    for i in range(n // 2): fn(i)
    e1.record(); torch.cuda.synchronize()
    print("   ", s)
    return e0.elapsed_time(e1) / n * 1000
A = lambda i: eng.step(**kw)
B = lambda i: eng.step(**kw_enc)
big = torch.empty(1 << 29, dtype=torch.float32, device="cuda")
def reset():
    big.zero_(); eng.planes.zero_(); torch.cuda.synchronize()
out = []
reset()
for name, fn in (("A", A), ("A", A), ("E", e2e), ("A", A), ("reset", None), ("A", A), ("B", B), ("A", A), ("reset", None), ("A", A)):
    if fn is None:
        reset(); out.append("reset")
    else:
        out.append("%s %.1f" % (name, t(fn)))
print("hints=%s: %s" % (os.environ.get("ASZ_ENV_HINTS", "default"), " | ".join(out)))
