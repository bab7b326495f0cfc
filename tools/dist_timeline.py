"""Timeline helper (not a test): per-kernel start offsets / durations of ONE distributed NT-Xent fwd+bwd step on rank 0,
taken with torch.profiler (CUPTI) while every rank runs the same loop.  torchrun, N ranks.  Shows where the time
between the two tensor-core kernels goes (pushes, flag waits, copies, memsets, launch gaps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import torch, torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from ssv_b200.dist import DistributedSimclrLoss

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)
N, d = int(os.environ.get("N", 32768)), 128
n = N // world
g = torch.Generator().manual_seed(rank)
zi = torch.randn(n, d, generator=g).to(dev).requires_grad_(True)
zj = torch.randn(n, d, generator=g).to(dev).requires_grad_(True)
fn = DistributedSimclrLoss(True, 0.5, transport=os.environ.get("TRANSPORT", "auto"))


def step():
    zi.grad = None; zj.grad = None
    fn(zi, zj).backward()


for _ in range(10):
    step()
dist.barrier(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(6):
        step()
    torch.cuda.synchronize()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    # split into steps at every pair_prep / p2p_prep_push kernel; print the 4th step
    starts = [i for i, e in enumerate(evs) if "prep" in e.name]
    if len(starts) >= 5:
        a, b = starts[3], starts[4]
        t0 = evs[a].time_range.start
        print(f"world={world} N={N}: one step on rank 0 ({(evs[b].time_range.start - t0):.1f} us from prep to next prep)")
        prev_end = t0
        for e in evs[a:b]:
            s, dur = e.time_range.start - t0, e.time_range.end - e.time_range.start
            print(f"  +{s:8.1f} us  gap {e.time_range.start - prev_end:6.1f}  dur {dur:8.1f}  {e.name[:90]}")
            prev_end = e.time_range.end
dist.barrier()
dist.destroy_process_group()
