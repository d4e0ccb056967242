"""Experiment for the two timing regimes of env_step_kernel: does a READ sweep over more than the L2's capacity
(which forces every dirty line out without adding new dirty lines) bring the fast regime back, and what sends it away?
A = tic + encode (the bench workload), B = encode only, sweepN = sum() over N MB of an unrelated tensor, fill = fill_ of 1 GB."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from alphasnake_zero_b200.engine import Engine  # noqa: E402

G = 65536
eng = Engine(side=11, snakes=4, health_dec=1, games=G, seed=1)
eng.reset()
_ = eng.planes
KW = dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True)
KW_ENC = dict(tic=False, encode=True)
KW_TIC = dict(spawn_mode=2, tic=True, encode=False, auto_reset=True, random_actions=True)
src = torch.ones(1 << 28, dtype=torch.float32, device="cuda")        # 1 GB, written once, long before it is read
big = torch.empty(1 << 28, dtype=torch.float32, device="cuda")


try:
    import pynvml
    pynvml.nvmlInit()
    NV = pynvml.nvmlDeviceGetHandleByIndex(0)
except Exception:
    NV = None


def smi():
    if NV is None:
        return ""
    try:
        return " [sm %d MHz, mem %d MHz, %.0f W, reasons 0x%x]" % (
            pynvml.nvmlDeviceGetClockInfo(NV, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(NV, pynvml.NVML_CLOCK_MEM),
            pynvml.nvmlDeviceGetPowerUsage(NV) / 1000.0, pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(NV))
    except Exception as ex:
        return " [nvml: %r]" % ex


def t(kw, n=300):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.step(**kw)
    info = smi()          # the launches are queued and executing
    e1.record()
    torch.cuda.synchronize()
    return "%.1f%s" % (e0.elapsed_time(e1) / n * 1000, info)


def sweep(mb):
    n = mb * (1 << 18)
    s = src[:n].sum()
    torch.cuda.synchronize()
    return float(s)


out = []
for _ in range(20):
    eng.step(**KW)
seq = sys.argv[1].split(",") if len(sys.argv) > 1 else \
    "A,A,sweep256,A,A,B,A,A,sweep256,A,A,fill,A,A,sweep128,A,B,A,sweep1024,A,A,tic,A,idle,A,B,A,idle,A,sweep256,A".split(",")
for op in seq:
    if op == "A":
        out.append("A %s" % t(KW, 600))
    elif op == "B":
        out.append("B %s" % t(KW_ENC, 100))
    elif op == "tic":
        out.append("tic %s" % t(KW_TIC, 100))
    elif op == "fill":
        big.fill_(2.0); torch.cuda.synchronize(); out.append("fill")
    elif op == "idle":
        time.sleep(0.5); out.append("idle")
    elif op == "cond":
        eng.condition_l2(); torch.cuda.synchronize(); out.append("cond")
    elif op == "condns":           # no synchronisation after the sweep: the next launches queue right behind it
        eng.condition_l2(); out.append("condns")
    elif op == "fillns":
        big.fill_(2.0); out.append("fillns")
    elif op == "Bns":
        for _ in range(100):
            eng.step(**KW_ENC)
        out.append("Bns")
    elif op == "Ans":
        for _ in range(300):
            eng.step(**KW)
        out.append("Ans")
    elif op == "sync":
        torch.cuda.synchronize(); out.append("sync")
    elif op == "dirty":            # what asz_reset does to the engine's flag, without touching anything else
        eng.step(tic=False, encode=True); out.append("dirty")
    elif op.startswith("sweep"):
        sweep(int(op[5:])); out.append(op)
print("   L2 monitor: %s" % {k: v for k, v in eng.totals().items() if k.startswith("l2_")})
print(os.environ.get("ASZ_LIB", "default"), "hints=%s" % os.environ.get("ASZ_ENV_HINTS", "1"), "|", " | ".join(out))
