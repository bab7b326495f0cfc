"""Timing helper (not a test): host (Python + launch) time per fwd+bwd step of the loss modules, measured with a
problem so small that the GPU is never the bottleneck.  If this exceeds the device time of the real problem the
step is CPU-bound."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import torch, ssv_b200
from ssv_b200.dist import DistributedSimclrLoss, DistributedBarlowLoss, DistributedSwavLoss, DistributedMocoLoss

dev = torch.device("cuda", 0)
def bench(name, fn, n=300):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name:40s} host {1e6 * (t1 - t0) / n:7.1f} us/step")

a = torch.randn(256, 128, device=dev, requires_grad=True); b = torch.randn(256, 128, device=dev, requires_grad=True)
def step(fn, *t):
    def run():
        for x in t: x.grad = None
        fn().backward()
    return run
f1 = ssv_b200.SimclrLoss(True, 0.5); bench("SimclrLoss", step(lambda: f1(a, b), a, b))
f2 = DistributedSimclrLoss(True, 0.5); bench("DistributedSimclrLoss (world 1)", step(lambda: f2(a, b), a, b))
f3 = DistributedBarlowLoss(False, 0.005); bench("DistributedBarlowLoss (world 1)", step(lambda: f3(a, b), a, b))
f3b = ssv_b200.BarlowLoss(False, 0.005); bench("BarlowLoss", step(lambda: f3b(a, b), a, b))
pc = torch.nn.functional.normalize(torch.randn(64, 128, device=dev)).requires_grad_(True)
f4 = DistributedSwavLoss(); bench("DistributedSwavLoss (world 1)", step(lambda: f4(a, b, pc), a, b, pc))
f4b = ssv_b200.SwavLoss(); bench("SwavLoss", step(lambda: f4b(a, b, pc), a, b, pc))
q = torch.nn.functional.normalize(torch.randn(1024, 128, device=dev))
f5 = DistributedMocoLoss(True, 0.07); bench("DistributedMocoLoss (world 1)", step(lambda: f5(a, b, q), a, b))
f5b = ssv_b200.MocoLoss(True, 0.07); bench("MocoLoss", step(lambda: f5b(a, b, q), a, b))

# ---- where the host time goes (cProfile of 2000 SimclrLoss fwd+bwd steps; cumulative, top 35)
import cProfile, pstats, io
run = step(lambda: f1(a, b), a, b)
pr = cProfile.Profile()
pr.enable()
for _ in range(2000):
    run()
pr.disable()
torch.cuda.synchronize()
sio = io.StringIO()
pstats.Stats(pr, stream=sio).sort_stats("cumulative").print_stats(35)
print(sio.getvalue()[:6000])
# forward only / backward only split
import time
def fwd_only():
    return f1(a, b)
for _ in range(50): fwd_only()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(1000): l = fwd_only()
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"SimclrLoss forward only: host {1e6 * (t1 - t0) / 1000:.1f} us")
from ssv_b200 import _cabi as C
L = C.lib()
t0 = time.perf_counter()
for _ in range(20000): L.ssvb_version()
print(f"bare ctypes call: {1e6 * (time.perf_counter() - t0) / 20000:.2f} us")
t0 = time.perf_counter()
for _ in range(20000): C.stream_ptr(dev)
print(f"stream_ptr: {1e6 * (time.perf_counter() - t0) / 20000:.2f} us")
t0 = time.perf_counter()
for _ in range(20000): torch.empty((), dtype=torch.float32, device=dev)
print(f"torch.empty scalar: {1e6 * (time.perf_counter() - t0) / 20000:.2f} us")
t0 = time.perf_counter()
for _ in range(20000): C.ptr(a)
print(f"C.ptr: {1e6 * (time.perf_counter() - t0) / 20000:.2f} us")
