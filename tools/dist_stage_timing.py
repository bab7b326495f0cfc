"""Timing helper (not a test): per-stage device time of the distributed NT-Xent step (torchrun, N ranks), for the
NCCL transport and the peer-memory (symmetric memory) transport."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
import torch, torch.distributed as dist
from ssv_b200.dist import CudaStages, _gather_slots, _PeerArena

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
colstat = torch.empty(mpad, device=dev); loss = torch.empty((), device=dev)
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


def make_p2p_body(arena):
    def p2p_body(ev):
        gen = arena.next_gen()
        st.p2p_prep_push(zi, zj, 1, tau, world, rank, arena, gen, inv, pos); ev[1].record()
        st.p2p_wait_copy(arena, world, rank, n, d, gen, zhat); ev[2].record()
        st.p2p_rows_fwd(zhat, world, rank, n, d, 1, tau, pos, arena, gen, ls); ev[3].record()
        st.p2p_stat_loss(arena, world, rank, n, d, 1, tau, gen, colstat, loss); ev[4].record()
        st.p2p_rows_bwd(zi, zj, 1, tau, world, rank, zhat, colstat, inv, go, dzi, dzj); ev[5].record()
    return p2p_body


if rank == 0: print("--- NCCL transport")
run(["prep", "allgather zhat", "rows_fwd", "allgather stat", "loss", "rows_bwd"], nccl_body)
P2P = ["prep+push+flags", "wait+copy", "rows_fwd+push+flags", "wait+stat+loss", "rows_bwd"]
if world > 1:
    for mc in (True, False):
        arena = _PeerArena.get(None, world, n, d, dev, multicast=mc)
        if rank == 0: print(f"--- peer-memory transport, multicast={'on' if (mc and arena.multicast_ptr) else 'off'}")
        run(P2P, make_p2p_body(arena))
dist.destroy_process_group()
