"""Timing helper (not a test): per-stage device time of the distributed NT-Xent step (torchrun, N ranks), for the
NCCL transport and the peer-memory (symmetric memory) transport."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import torch, torch.distributed as dist
from ssv_b200.dist import CudaStages, _gather_slots, _PeerTransport

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)
N, d, tau = 32768, 128, 0.5
n = N // world
st = CudaStages()
g = torch.Generator().manual_seed(rank)
zi = torch.randn(n, d, generator=g).to(dev); zj = torch.randn(n, d, generator=g).to(dev)
m = 2 * N
mpad, dpad = st.mpad(N), st.dpad(d)
zhat = torch.empty(mpad, dpad, dtype=torch.bfloat16, device=dev)
inv = torch.empty(2 * n, device=dev); pos = torch.empty(2 * n, device=dev)
stat = torch.empty(world, 2, 2 * n, device=dev); ls = torch.zeros((), device=dev)
go = torch.ones((), device=dev); dzi = torch.empty_like(zi); dzj = torch.empty_like(zj)
my = slice(rank * 2 * n, (rank + 1) * 2 * n)
peer = _PeerTransport.get(None, world, mpad, dpad, n, dev)
reps = 30


def run(names, body):
    acc = [0.0] * len(names)
    for it in range(reps + 5):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        dist.barrier(); torch.cuda.synchronize()
        ev[0].record()
        body(ev)
        torch.cuda.synchronize()
        if it >= 5:
            for i in range(len(names)):
                acc[i] += ev[i].elapsed_time(ev[i + 1])
    if rank == 0:
        print(f"world={world}: total {sum(acc) / reps * 1e3:.0f} us per step (stage-serialised device time)")
        for nm, a in zip(names, acc):
            print(f"  {nm:18s} {a / reps * 1e3:8.1f} us")


def nccl_body(ev):
    st.prep(zi, zj, 1, 0.5, world, rank, zhat, inv, pos); ev[1].record()
    _gather_slots(zhat[:m], zhat[my], None, True); ev[2].record()
    st.rows_fwd(zhat, world, rank, n, d, 1, tau, pos, stat[rank], ls); ev[3].record()
    _gather_slots(stat.view(world * 2, 2 * n), stat[rank], None, True); ev[4].record()
    st.dist_loss(stat, world, n, ls); ev[5].record()
    st.rows_bwd(zi, zj, 1, tau, world, rank, zhat, stat, inv, go, dzi, dzj); ev[6].record()


def p2p_body(ev):
    zbuf, hz, sbuf, hs = peer.next()
    st.prep_push(zi, zj, 1, 0.5, world, rank, hz.buffer_ptrs_dev, inv, pos); ev[1].record()
    hz.barrier(); ev[2].record()
    z = zbuf.view(mpad, dpad).clone()
    st.rows_fwd_push(z, world, rank, n, d, 1, tau, pos, hs.buffer_ptrs_dev, ls); ev[3].record()
    hs.barrier(); ev[4].record()
    s_all = sbuf.view(world, 2, 2 * n); st.dist_loss(s_all, world, n, ls); ev[5].record()
    st.rows_bwd(zi, zj, 1, tau, world, rank, z, s_all, inv, go, dzi, dzj); ev[6].record()


if rank == 0: print("--- NCCL transport")
run(["prep", "allgather zhat", "rows_fwd", "allgather stat", "loss", "rows_bwd"], nccl_body)
if rank == 0: print("--- peer-memory transport")
run(["prep+push", "barrier", "rows_fwd+push", "barrier", "loss", "rows_bwd"], p2p_body)
dist.destroy_process_group()
