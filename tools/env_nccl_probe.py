"""What does initialising an NCCL communicator change for env_step_kernel?  Prints the device limits before / after
dist.init_process_group('nccl') and times the kernel; optional fixes by name: resetl2, stack, destroy."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from alphasnake_zero_b200.engine import Engine  # noqa: E402

fixes = sys.argv[1].split(",") if len(sys.argv) > 1 else []
rt = C.CDLL("libcudart.so.12")
LIMITS = {"stack": 0, "printf_fifo": 1, "malloc_heap": 2, "sync_depth": 3, "pending_launch": 4, "max_l2_fetch_granularity": 5, "persisting_l2": 6}


def limits():
    out = {}
    for k, v in LIMITS.items():
        x = C.c_size_t(0)
        rc = rt.cudaDeviceGetLimit(C.byref(x), v)
        out[k] = x.value if rc == 0 else "rc%d" % rc
    rt.cudaGetLastError()
    return out


torch.cuda.set_device(0)
torch.zeros(1, device="cuda")
print("limits before:", limits())
if "nonccl" not in fixes:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0)) if "lazy" not in fixes else \
        dist.init_process_group("nccl", rank=0, world_size=1)
    if "nobarrier" not in fixes:
        dist.barrier()
    torch.cuda.synchronize()
print("limits after :", limits())
if "destroy" in fixes:
    dist.destroy_process_group()
if "stack" in fixes:
    print("set stack 1024:", rt.cudaDeviceSetLimit(0, C.c_size_t(1024)))
if "resetl2" in fixes:
    print("reset persisting l2:", rt.cudaCtxResetPersistingL2Cache(), rt.cudaDeviceSetLimit(6, C.c_size_t(0)))
if "fetch" in fixes:
    print("set fetch granularity 128:", rt.cudaDeviceSetLimit(5, C.c_size_t(128)))
G = 65536
eng = Engine(side=11, snakes=4, health_dec=1, games=G, seed=1)
eng.reset()
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
print(fixes, "|", " ".join("%.1f" % t() for _ in range(3)), "us |", {k: v for k, v in eng.totals().items() if k.startswith("l2_")})
