"""The two timing regimes of env_step_kernel, one script, four experiments:

  python tools/env_regimes.py ncu      device path, e2e path, device path again: meant to run UNDER ncu (single-pass metrics)
  python tools/env_regimes.py events   per-step GPU time (CUDA events) next to the CPU enqueue time, phases A / B / C / D
  python tools/env_regimes.py phases   phase timings from a "reset" start (streaming writes); ASZ_ENV_HINTS / ASZ_ENV_HINTS_HOST
  python tools/env_regimes.py fill     A (tic + encode), B (encode only), C (A again), then C with a streaming fill before every launch

A = device-resident tic + encode, B = encode only, E = asz_env_step_host (host buffers), all at configs[1] size."""
import ctypes as C
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from alphasnake_zero_b200 import _lib  # noqa: E402
from alphasnake_zero_b200.engine import Engine  # noqa: E402

G = 65536
KW = dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True)
KW_ENC = dict(tic=False, encode=True)


def make():
    eng = Engine(side=11, snakes=4, health_dec=1, games=G, seed=1)
    eng.reset()
    _ = eng.planes
    return eng


def host_stepper(eng):
    rng = np.random.default_rng(0)
    pool = [torch.from_numpy(rng.integers(0, 3, size=(G, 8), dtype=np.uint8)).pin_memory() for _ in range(8)]
    h_ended = torch.zeros(G, dtype=torch.uint8).pin_memory()
    h_rewards = torch.zeros(G, 8, dtype=torch.int8).pin_memory()
    rows = C.c_int32(0)
    L = _lib.lib()
    flags = _lib.STEP_TIC | _lib.STEP_ENCODE | _lib.STEP_AUTO_RESET

    def step(i):
        _lib.check(L.asz_env_step_host(eng.h, flags, _lib.SPAWN_NATIVE, C.c_void_p(pool[i % 8].data_ptr()), None,
                                       C.c_void_p(h_ended.data_ptr()), C.c_void_p(h_rewards.data_ptr()), C.byref(rows), None, None,
                                       eng.stream))
    return step


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000


def mode_ncu():
    eng = make()
    e2e = host_stepper(eng)
    for i in range(30):
        eng.step(**KW)        # launches 0..29: device path
    torch.cuda.synchronize()
    for i in range(30):
        e2e(i)                # launches 30..59: host-buffer path
    for i in range(30):
        eng.step(**KW)        # launches 60..89: device path again
    torch.cuda.synchronize()


def mode_events():
    eng = make()

    def probe(name, kw, n=200):
        for _ in range(20):
            eng.step(**kw)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
        t0 = time.perf_counter()
        ev[0].record()
        for i in range(n):
            eng.step(**kw)
            ev[i + 1].record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        d = np.array([ev[i].elapsed_time(ev[i + 1]) * 1000 for i in range(n)])
        print("%-28s cpu enqueue %.1f us/step, wall %.1f us/step | gpu per step: median %.1f min %.1f p90 %.1f max %.1f" %
              (name, (t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6, np.median(d), d.min(), np.percentile(d, 90), d.max()))
    probe("A tic+encode", KW)
    probe("B encode only", KW_ENC)
    probe("C tic+encode", KW)
    probe("C tic+encode", KW)
    eng.planes.zero_()
    probe("D tic+encode after zero_", KW)


def mode_phases():
    eng = make()
    e2e = host_stepper(eng)

    def smi():
        q = "clocks.sm,clocks.gr,clocks.mem,clocks.video,power.draw,pstate"
        return subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=" + q, "--format=csv,noheader"], capture_output=True,
                              text=True).stdout.strip()
    big = torch.empty(1 << 29, dtype=torch.float32, device="cuda")

    def reset():
        big.zero_(); eng.planes.zero_(); torch.cuda.synchronize()
    A = lambda i: eng.step(**KW)          # noqa: E731
    B = lambda i: eng.step(**KW_ENC)      # noqa: E731
    out = []
    reset()
    for name, fn in (("A", A), ("A", A), ("E", e2e), ("A", A), ("reset", None), ("A", A), ("B", B), ("A", A), ("reset", None), ("A", A)):
        if fn is None:
            reset(); out.append("reset")
        else:
            out.append("%s %.1f" % (name, timed(fn, 600)))
            print("   ", smi())
    print("hints=%s: %s" % (os.environ.get("ASZ_ENV_HINTS", "default"), " | ".join(out)))


def mode_fill():
    eng = make()
    for _ in range(20):
        eng.step(**KW)
    t = lambda kw, n=300, pre=None: timed(lambda i: ((pre() if pre else None), eng.step(**kw)), n)   # noqa: E731
    print("A", " ".join("%.1f" % t(KW) for _ in range(2)))
    print("B", "%.1f" % t(KW_ENC))
    print("C", " ".join("%.1f" % t(KW) for _ in range(2)))
    for mb in (16, 64, 256):
        x = torch.empty(mb * (1 << 18), dtype=torch.float32, device="cuda")
        print("C with a %d MB fill before every launch:" % mb, " ".join("%.1f" % t(KW, 300, lambda: x.fill_(0.0)) for _ in range(2)),
              " then plain C:", " ".join("%.1f" % t(KW) for _ in range(2)))
        print("B", "%.1f" % t(KW_ENC), " C", " ".join("%.1f" % t(KW) for _ in range(2)))


if __name__ == "__main__":
    {"ncu": mode_ncu, "events": mode_events, "phases": mode_phases, "fill": mode_fill}[sys.argv[1] if len(sys.argv) > 1 else "events"]()
