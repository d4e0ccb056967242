"""Does the env kernel hold its rate over seconds, and what perturbs it?  Per-300-launch timings."""
import os, sys, subprocess, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from alphasnake_zero_b200.engine import Engine
eng = Engine(side=11, snakes=4, health_dec=1, games=65536, seed=1)
eng.reset(); _ = eng.planes
kw = dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True)
kw_enc = dict(tic=False, encode=True)
kw_tic = dict(spawn_mode=2, tic=True, encode=False, auto_reset=True, random_actions=True)
def t(kw, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): eng.step(**kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
def live():
    return float(eng.alive_mask().float().sum().item()) / 65536
def phase(name, kw, reps, n=300):
    print(name, " ".join("%.1f" % t(kw, n) for _ in range(reps)), "us | live/game %.3f rows %d" % (live(), int(eng.row_count.item())))
for _ in range(20): eng.step(**kw)
torch.cuda.synchronize()
def smi():
    q = "clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active"
    return subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=" + q, "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
x = torch.empty(993462041 // 4, dtype=torch.float32, device="cuda")
def fill_us(n=50):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): x.fill_(1.0)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
phase("A tic+encode        ", kw, 3)
print("fill_ %.1f us, tic only %.1f us" % (fill_us(), t(kw_tic, 100)))
phase("B encode only x300  ", kw_enc, 1)
phase("C tic+encode        ", kw, 2)
print("fill_ %.1f us, tic only %.1f us" % (fill_us(), t(kw_tic, 100)))
phase("C tic+encode        ", kw, 2)
st = eng.get_state(0)
phase("C after get_state (device sync + small copies)", kw, 2)
eng.reset()
for _ in range(40): eng.step(**kw_tic)
phase("C after reset + 40 tics", kw, 3)
phase("B encode only x300  ", kw_enc, 1)
phase("C tic+encode        ", kw, 2)
eng2 = Engine(side=11, snakes=4, health_dec=1, games=65536, seed=2)
eng2.reset(); _ = eng2.planes
for _ in range(40): eng2.step(**kw_tic)
def t2(kw, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): eng2.step(**kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
print("second engine tic+encode", " ".join("%.1f" % t2(kw, 300) for _ in range(3)))
phase("C first engine again", kw, 2)
