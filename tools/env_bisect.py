"""bisect what puts a fresh process into the persistent slow regime: flags = comma list of: dist (import torch.distributed),
cudart (dlopen libcudart + cudaDeviceGetLimit calls), zeros (a tiny torch allocation first), planes (touch eng.planes before stepping),
big (allocate 4 GB with torch first), gloo (init a gloo group)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

flags = sys.argv[1].split(",") if len(sys.argv) > 1 and sys.argv[1] else []
if "dist" in flags or "gloo" in flags or "nccl" in flags:
    import torch.distributed as dist
from alphasnake_zero_b200.engine import Engine  # noqa: E402

torch.cuda.set_device(0)
if "zeros" in flags:
    torch.zeros(1, device="cuda")
if "cudart" in flags:
    rt = C.CDLL("libcudart.so.12")
    for v in range(7):
        x = C.c_size_t(0)
        rt.cudaDeviceGetLimit(C.byref(x), v)
    rt.cudaGetLastError()
for f in flags:
    if f.startswith("big"):          # bigN: N x 256 MB allocated (and kept) before the engine
        keep = torch.empty(int(f[3:] or 16) << 26, dtype=torch.float32, device="cuda")
    if f.startswith("freed"):        # freedN: allocated and released again before the engine
        tmp = torch.empty(int(f[5:] or 16) << 26, dtype=torch.float32, device="cuda"); del tmp; torch.cuda.empty_cache()
if "gloo" in flags or "nccl" in flags:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29534")
    dist.init_process_group("gloo" if "gloo" in flags else "nccl", rank=0, world_size=1)
    if "nccl" in flags:
        dist.barrier(); torch.cuda.synchronize()
eng = Engine(side=11, snakes=4, health_dec=1, games=65536, seed=1)
eng.reset()
for f in flags:
    if f.startswith("mid"):          # midN: N x 256 MB allocated after the engine, before the batch buffer
        keep2 = torch.empty(int(f[3:] or 16) << 26, dtype=torch.float32, device="cuda")
if "planes" in flags:
    _ = eng.planes
for f in flags:
    if f.startswith("late"):         # lateN: allocated after everything the kernel touches
        _ = eng.planes
        keep3 = torch.empty(int(f[4:] or 16) << 26, dtype=torch.float32, device="cuda")
dense = None
if "dense" in flags:
    dense = torch.empty(65536 * 4, 21, 21, 3, device="cuda")
kw = dict(spawn_mode=2, tic=True, encode=True, auto_reset=True, random_actions=True, planes=dense)


def t(n=300):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.step(**kw)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000


for _ in range(50):
    eng.step(**kw)
print("%-28s" % ",".join(flags), "|", " ".join("%.1f" % t() for _ in range(3)), "us |", {k: v for k, v in eng.totals().items() if k.startswith("l2_")},
      "| planes ptr %x" % eng.planes.data_ptr())
