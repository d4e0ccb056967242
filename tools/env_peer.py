"""Does peer access (what NCCL enables between the GPUs of a node) change the regime of env_step_kernel on ONE GPU?
  python tools/env_peer.py none|before|after     (2-GPU box; one process, engine on cuda:0)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from alphasnake_zero_b200.engine import Engine  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "none"
rt = C.CDLL("libcudart.so.12")


def enable_peer():
    torch.cuda.set_device(0)
    r0 = rt.cudaDeviceEnablePeerAccess(1, 0)
    torch.cuda.set_device(1)
    r1 = rt.cudaDeviceEnablePeerAccess(0, 0)
    torch.cuda.set_device(0)
    rt.cudaGetLastError()
    return r0, r1


torch.cuda.set_device(0)
torch.zeros(1, device="cuda:0"); torch.zeros(1, device="cuda:1")
msg = ""
if mode == "before":
    msg = "peer access enabled BEFORE the engine's allocations: rc %s" % (enable_peer(),)
G = 65536
eng = Engine(side=11, snakes=4, health_dec=1, games=G, seed=1)
eng.reset()
_ = eng.planes
if mode == "after":
    msg = "peer access enabled AFTER the engine's allocations: rc %s" % (enable_peer(),)
kw = dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True)


def t(n=400):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.step(**kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000


for _ in range(50):
    eng.step(**kw)
print(mode, "|", msg, "|", " ".join("%.1f" % t() for _ in range(3)), "us |", {k: v for k, v in eng.totals().items() if k.startswith("l2_")})
