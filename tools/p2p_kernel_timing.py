"""Timing helper (not a test): pure device time of each kernel of the NT-Xent peer-memory transport, measured with
CUDA events around REPS back-to-back launches of the same stage (no host gaps, no cross-stage waits except the ones
the protocol itself contains).  torchrun, N ranks."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import torch, torch.distributed as dist
from ssv_b200.dist import CudaStages, _PeerArena

rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr_)
dev = torch.device("cuda", lr_)
dist.init_process_group("nccl", device_id=dev)
N, d, tau = int(os.environ.get("N", 32768)), 128, 0.5
n = N // world
st = CudaStages()
g = torch.Generator().manual_seed(rank)
zi = torch.randn(n, d, generator=g).to(dev); zj = torch.randn(n, d, generator=g).to(dev)
mpad, dpad = st.mpad(N), st.dpad(d)
zhat = torch.empty(mpad, dpad, dtype=torch.bfloat16, device=dev)
inv = torch.empty(2 * n, device=dev); pos = torch.empty(2 * n, device=dev)
ls = torch.zeros((), device=dev); loss = torch.empty((), device=dev); colstat = torch.empty(mpad, device=dev)
go = torch.ones((), device=dev); dzi = torch.empty_like(zi); dzj = torch.empty_like(zj)
REPS = 10


def timed(name, fn, sync_ranks=True):
    if sync_ranks:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(REPS):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / REPS * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"  {name:44s} {t.item():8.1f} us / launch (max over ranks)")


for mc in (True, False):
    arena = _PeerArena.get(None, world, n, d, dev, multicast=mc)
    if rank == 0:
        print(f"--- world={world} N={N}: multicast={'on' if (mc and arena.multicast_ptr) else 'off'}")
    # every stage below is protocol-safe to repeat: generations advance by one full step per repetition
    base = arena.gen

    def full_step(i):
        gen = arena.next_gen()
        st.p2p_prep_push(zi, zj, 1, tau, world, rank, arena, gen, inv, pos)
        st.p2p_wait_copy(arena, world, rank, n, d, gen, zhat)
        st.p2p_rows_fwd(zhat, world, rank, n, d, 1, tau, pos, arena, gen, ls)
        st.p2p_stat_loss(arena, world, rank, n, d, 1, tau, gen, colstat, loss)
        st.p2p_rows_bwd(zi, zj, 1, tau, world, rank, zhat, colstat, inv, go, dzi, dzj)
    for _ in range(3):
        full_step(0)
    timed("full step (5 stages back to back)", full_step)

    def exchange_only(i):   # the two exchanges without the tensor-core kernels: push, wait+copy; statistics are stale but flagged
        gen = arena.next_gen()
        st.p2p_prep_push(zi, zj, 1, tau, world, rank, arena, gen, inv, pos)
        st.p2p_wait_copy(arena, world, rank, n, d, gen, zhat)
        # keep the protocol's second exchange alive so generations stay aligned on every rank
        st.p2p_rows_fwd(zhat, world, rank, n, d, 1, tau, pos, arena, gen, ls)
        st.p2p_stat_loss(arena, world, rank, n, d, 1, tau, gen, colstat, loss)
    timed("forward only (push, wait+copy, fwd, stat+loss)", exchange_only)
    # re-pushing an already consumed generation is harmless while every rank is in this same loop (nobody reads)
    timed("prep_push alone (push + fence + flags)", lambda i: st.p2p_prep_push(zi, zj, 1, tau, world, rank, arena, arena.gen, inv, pos))
    timed("wait_copy alone (flags already set)", lambda i: st.p2p_wait_copy(arena, world, rank, n, d, arena.gen, zhat))
    timed("stat_loss alone (flags already set)", lambda i: st.p2p_stat_loss(arena, world, rank, n, d, 1, tau, arena.gen, colstat, loss))
    timed("rows_bwd only", lambda i: st.p2p_rows_bwd(zi, zj, 1, tau, world, rank, zhat, colstat, inv, go, dzi, dzj))
    timed("plain prep (no push), NCCL-path kernel", lambda i: st.prep(zi, zj, 1, tau, world, rank, zhat, inv, pos))
dist.destroy_process_group()
