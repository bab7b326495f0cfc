"""Multi-GPU parity (needs >= 2 B200s; skipped on a single-GPU box): row-sharded global-batch NT-Xent over NCCL
vs the single-process oracle on the concatenated batch, and vs the single-GPU CUDA path."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_local, d, normalize, tau, out, transport="auto"):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from ssv_b200.dist import DistributedSimclrLoss
    g = torch.Generator().manual_seed(100 + rank)
    zi = torch.randn(n_local, d, generator=g)
    zj = torch.randn(n_local, d, generator=g)
    a = zi.cuda().requires_grad_(True)
    b = zj.cuda().requires_grad_(True)
    fn = DistributedSimclrLoss(normalize, tau, transport=transport)
    for _ in range(3):  # several steps: exercises the double-buffered peer transport
        a.grad = None; b.grad = None
        loss = fn(a, b)
        loss.backward()
    torch.cuda.synchronize()
    out[rank] = (loss.item(), a.grad.cpu().numpy(), b.grad.cpu().numpy(), zi.numpy(), zj.numpy())
    if transport.startswith("p2p"):
        # the backward only reads private copies: any forward / backward interleaving is valid (several forwards in
        # flight, both buffer parities overwritten in between)
        l1 = fn(a, b); l2 = fn(a, b); l3 = fn(a, b)
        a.grad = None; b.grad = None
        l1.backward()
        torch.cuda.synchronize()
        g_ = a.grad.cpu().numpy()  # (column-chunked accumulation uses red.add: equal up to fp32 summation order)
        assert np.linalg.norm(g_ - out[rank][1]) <= 1e-5 * np.linalg.norm(out[rank][1])
        assert l1.item() == l2.item() == l3.item() == out[rank][0]
        del l2, l3
        # the loss is bit-identical on every rank (same gathered terms, same fixed-order reduction)
        t = torch.tensor([out[rank][0]], dtype=torch.float64, device="cuda")
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert lo.item() == hi.item()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["p2p", "p2p-unicast", "nccl"])
@pytest.mark.parametrize("n_local,d,normalize,tau", [(192, 128, True, 0.5), (1000, 64, True, 0.07), (256, 128, True, 0.02)])
def test_dist_ntxent_vs_oracle(n_local, d, normalize, tau, transport):
    _dist_ntxent_case(n_local, d, normalize, tau, transport)


def test_dist_ntxent_wide_rows_take_nccl():
    """128 < d <= 256: the four-k-block kernels behind the NCCL transport (the peer-push kernels cover d <= 128, so
    transport="auto" must fall back without being asked)."""
    _dist_ntxent_case(160, 200, True, 0.5, "auto")


def _dist_ntxent_case(n_local, d, normalize, tau, transport):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import ssl_oracle as O
    world = min(torch.cuda.device_count(), 4)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29700 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_local, d, normalize, tau, out, transport), nprocs=world, join=True)
    zi = np.concatenate([out[r][3] for r in range(world)])
    zj = np.concatenate([out[r][4] for r in range(world)])
    ref_loss, ref_dzi, ref_dzj = O.ntxent(zi, zj, normalize, tau)
    for r in range(world):
        loss, gi, gj, _, _ = out[r]
        assert abs(loss - ref_loss) / abs(ref_loss) <= 1e-3
        sl = slice(r * n_local, (r + 1) * n_local)
        assert np.linalg.norm(gi - ref_dzi[sl]) / np.linalg.norm(ref_dzi[sl]) <= 1e-2
        assert np.linalg.norm(gj - ref_dzj[sl]) / np.linalg.norm(ref_dzj[sl]) <= 1e-2


# ======================================================================================================= other losses
def _init(rank, world, port):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))


def _unit(x):
    return torch.nn.functional.normalize(x, dim=1)


def _barlow_worker(rank, world, port, n_local, d, normalize, out, mode="allreduce"):
    _init(rank, world, port)
    from ssv_b200.dist import DistributedBarlowLoss
    g = torch.Generator().manual_seed(200 + rank)
    zi = torch.randn(n_local, d, generator=g) * 1.5 + 0.3
    zj = zi * 0.7 + 0.5 * torch.randn(n_local, d, generator=g)
    a, b = zi.cuda().requires_grad_(True), zj.cuda().requires_grad_(True)
    fn = DistributedBarlowLoss(normalize, 0.005, mode=mode)
    for _ in range(2):
        a.grad = None; b.grad = None
        loss = fn(a, b)
        loss.backward()
    torch.cuda.synchronize()
    out[rank] = (loss.item(), a.grad.cpu().numpy(), b.grad.cpu().numpy(), zi.numpy(), zj.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_local,d,normalize,mode", [(128, 1024, False, "allreduce"), (100, 264, True, "allreduce"),
                                                      (64, 1000, False, "allreduce"), (128, 1024, False, "colshard"),
                                                      (100, 256, True, "colshard")])
def test_dist_barlow_vs_oracle(n_local, d, normalize, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import ssl_oracle as O
    world = min(torch.cuda.device_count(), 4)
    if d == 1000:
        world = min(world, 3) if torch.cuda.device_count() >= 3 else 2  # 1000 % 3 != 0 -> all-reduce path at world 3
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29900 + (os.getpid() % 2000)
    mp.spawn(_barlow_worker, args=(world, port, n_local, d, normalize, out, mode), nprocs=world, join=True)
    zi = np.concatenate([out[r][3] for r in range(world)])
    zj = np.concatenate([out[r][4] for r in range(world)])
    ref_loss, ref_di, ref_dj = O.barlow(zi, zj, normalize, 0.005)
    for r in range(world):
        loss, gi, gj, _, _ = out[r]
        assert abs(loss - ref_loss) / abs(ref_loss) <= 1e-3
        sl = slice(r * n_local, (r + 1) * n_local)
        assert np.linalg.norm(gi - ref_di[sl]) / np.linalg.norm(ref_di[sl]) <= 1e-2
        assert np.linalg.norm(gj - ref_dj[sl]) / np.linalg.norm(ref_dj[sl]) <= 1e-2
    assert len({out[r][0] for r in range(world)}) == 1, "every rank must report the same global loss"


def _swav_worker(rank, world, port, nb, nbank, k, d, out):
    _init(rank, world, port)
    from ssv_b200.dist import DistributedSwavLoss
    g = torch.Generator().manual_seed(300 + rank)
    z1 = _unit(torch.randn(nb, d, generator=g))
    z2 = _unit(0.6 * z1 + 0.4 * torch.randn(nb, d, generator=g))
    bank = _unit(torch.randn(nbank, d, generator=g)) if nbank else None
    pc = _unit(torch.randn(k, d, generator=torch.Generator().manual_seed(7)))
    a, b, p = z1.cuda().requires_grad_(True), z2.cuda().requires_grad_(True), pc.cuda().requires_grad_(True)
    fn = DistributedSwavLoss(0.1, 0.05, 3)
    loss = fn(a, b, p, bank.cuda() if nbank else None)
    loss.backward()
    sc = (z1 @ pc.t()).contiguous()
    codes = fn.compute_codes_sinkhorn(sc.cuda())
    torch.cuda.synchronize()
    out[rank] = dict(loss=loss.item(), dz1=a.grad.cpu().numpy(), dz2=b.grad.cpu().numpy(), dpc=p.grad.cpu().numpy(),
                     z1=z1.numpy(), z2=z2.numpy(), bank=bank.numpy() if nbank else None, pc=pc.numpy(), sc=sc.numpy(),
                     codes=codes.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nb,nbank,k,d", [(256, 0, 3000, 128), (128, 200, 1000, 64)])
def test_dist_swav_vs_oracle(nb, nbank, k, d):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import ssl_oracle as O
    world = min(torch.cuda.device_count(), 4)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 30100 + (os.getpid() % 2000)
    mp.spawn(_swav_worker, args=(world, port, nb, nbank, k, d, out), nprocs=world, join=True)
    z1 = np.concatenate([out[r]["z1"] for r in range(world)])
    z2 = np.concatenate([out[r]["z2"] for r in range(world)])
    bank = np.concatenate([out[r]["bank"] for r in range(world)]) if nbank else None
    ref_loss, ref_dz1, ref_dz2, ref_dc = O.swav(z1, z2, out[0]["pc"], bank, 0.1, 0.05, 3)
    rl2 = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)  # noqa: E731
    for r in range(world):
        o = out[r]
        assert abs(o["loss"] - ref_loss) / abs(ref_loss) <= 1e-3
        sl = slice(r * nb, (r + 1) * nb)
        assert rl2(o["dz1"], ref_dz1[sl]) <= 1e-2 and rl2(o["dz2"], ref_dz2[sl]) <= 1e-2
    # prototype gradient: every rank returns its local contribution; their sum is the single-process gradient
    assert rl2(sum(out[r]["dpc"].astype(np.float64) for r in range(world)), ref_dc) <= 1e-2
    ref_codes = O.sinkhorn(np.concatenate([out[r]["sc"] for r in range(world)]), 0.05, 3)
    for r in range(world):
        assert rl2(out[r]["codes"], ref_codes[r * nb:(r + 1) * nb]) < 1e-4


def _max_ulp(a, b):
    """largest distance in units of the last place between two fp32 arrays (same sign pattern assumed; zeros exact)"""
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    assert ((a == 0) == (b == 0)).all() and (np.signbit(a) == np.signbit(b)).all()
    return int(np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64)).max())


def _moco_worker(rank, world, port, n_local, k_total, d, tau, out):
    _init(rank, world, port)
    from ssv_b200.dist import DistributedMocoLoss, ShardedMemoryBank
    g = torch.Generator().manual_seed(400 + rank)
    q = torch.randn(n_local, d, generator=g)
    k = torch.randn(n_local, d, generator=g)
    fill = torch.randn(k_total - 40, d, generator=torch.Generator().manual_seed(9))  # same on every rank
    bank = ShardedMemoryBank(k_total, d)
    per = fill.shape[0] // world  # each rank contributes its slice; the all-gather restores rank order
    bank.add_batch(fill[rank * per:(rank + 1) * per].cuda())
    a, b = q.cuda().requires_grad_(True), k.cuda().requires_grad_(True)
    before = bank.get_vectors().cpu().numpy().copy()
    loss = DistributedMocoLoss(True, tau)(a, b, bank.get_vectors())
    loss.backward()
    ptr_before = bank.ptr
    bank.add_batch(b.detach())   # wraps around the end of the ring
    torch.cuda.synchronize()
    out[rank] = dict(loss=loss.item(), dq=a.grad.cpu().numpy(), dk=b.grad.cpu().numpy(), q=q.numpy(), k=k.numpy(),
                     before=before, after=bank.get_vectors().cpu().numpy(), ptr_before=ptr_before, ptr=bank.ptr,
                     fill=fill[:per * world].numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_local,k_total,d,tau", [(64, 8192, 128, 0.07), (48, 1024, 64, 0.2)])
def test_dist_moco_sharded_queue_vs_oracle(n_local, k_total, d, tau):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import ssl_oracle as O
    world = 4 if torch.cuda.device_count() >= 4 else 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 30300 + (os.getpid() % 2000)
    mp.spawn(_moco_worker, args=(world, port, n_local, k_total, d, tau, out), nprocs=world, join=True)
    q = np.concatenate([out[r]["q"] for r in range(world)])
    k = np.concatenate([out[r]["k"] for r in range(world)])
    queue = np.concatenate([out[r]["before"] for r in range(world)])
    # the sharded ring after the fill == single-process ring fed the same global batch
    ref_bank, ref_ptr = O.ring_enqueue(np.zeros((k_total, d), np.float32), 0, out[0]["fill"], True)
    # normalised rows vs the fp64-rounded oracle: fp32 sum of squares (shuffle-tree order) + sqrt + division against a
    # correctly rounded reference -> 2 ulp typical, 3 ulp for ~1 element in 10^6 (measured on this test's 1 M elements)
    assert _max_ulp(queue, ref_bank) <= 3
    assert all(out[r]["ptr_before"] == ref_ptr for r in range(world))
    ref_loss, ref_dq, ref_dk = O.moco(q, k, queue, True, tau)
    rl2 = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)  # noqa: E731
    for r in range(world):
        o = out[r]
        assert abs(o["loss"] - ref_loss) / abs(ref_loss) <= 1e-3
        sl = slice(r * n_local, (r + 1) * n_local)
        assert rl2(o["dq"], ref_dq[sl]) <= 1e-2 and rl2(o["dk"], ref_dk[sl]) <= 1e-2
    assert len({out[r]["loss"] for r in range(world)}) == 1
    ref_bank2, ref_ptr2 = O.ring_enqueue(queue.copy(), ref_ptr, k, True)
    after = np.concatenate([out[r]["after"] for r in range(world)])
    assert _max_ulp(after, ref_bank2) <= 3
    assert all(out[r]["ptr"] == ref_ptr2 for r in range(world))


# ======================================================================================================= ReLIC + DINO centre
def _relic_worker(rank, world, port, n_local, d, tau, alpha, out):
    _init(rank, world, port)
    from ssv_b200.dist import DistributedRelicLoss, distributed_update_teacher_center
    g = torch.Generator().manual_seed(400 + rank)
    zi, zj, zo = (torch.randn(n_local, d, generator=g) for _ in range(3))
    a, b, c = (t.cuda().requires_grad_(True) for t in (zi, zj, zo))
    fn = DistributedRelicLoss(True, tau, alpha)
    for _ in range(2):
        a.grad = b.grad = c.grad = None
        loss = fn(a, b, c)
        loss.backward()
    teacher = torch.randn(n_local, 2, 256, generator=g)
    c0 = torch.linspace(-1, 1, 256)
    cen = distributed_update_teacher_center(c0.cuda(), teacher.cuda(), 0.9)
    torch.cuda.synchronize()
    out[rank] = (loss.item(), a.grad.cpu().numpy(), b.grad.cpu().numpy(), c.grad.cpu().numpy(), zi.numpy(), zj.numpy(),
                 zo.numpy(), cen.cpu().numpy(), teacher.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_local,d,tau,alpha", [(192, 128, 1.0, 0.5), (500, 64, 0.5, 2.0)])
def test_dist_relic_and_dino_center_vs_oracle(n_local, d, tau, alpha):
    """DistributedRelicLoss over NCCL / NVLink (reference utils/losses.py:154-201 on the concatenation; the KL's
    softmaxes span all ranks' rows) and the all-reduced DINO centre (models/dino.py:136-141)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from oracle import ssl_oracle as O
    world = min(torch.cuda.device_count(), 4)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 30300 + (os.getpid() % 2000)
    mp.spawn(_relic_worker, args=(world, port, n_local, d, tau, alpha, out), nprocs=world, join=True)
    zi, zj, zo = (np.concatenate([out[r][4 + k] for r in range(world)]) for k in range(3))
    ref = O.relic(zi, zj, zo, True, tau, alpha)
    teacher = np.concatenate([out[r][8] for r in range(world)]).reshape(-1, 256)
    ref_cen = O.dino_center_update(np.linspace(-1, 1, 256, dtype=np.float32), teacher, 0.9)
    for r in range(world):
        assert abs(out[r][0] - ref[0]) / abs(ref[0]) <= 1e-3
        sl = slice(r * n_local, (r + 1) * n_local)
        for k in range(3):
            assert np.linalg.norm(out[r][1 + k] - ref[1 + k][sl]) / np.linalg.norm(ref[1 + k][sl]) <= 1e-2
        np.testing.assert_allclose(out[r][7], ref_cen, rtol=1e-4, atol=1e-6)
        assert np.array_equal(out[r][7], out[0][7]), "every rank must hold the identical centre"
