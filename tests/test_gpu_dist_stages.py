"""Single-GPU parity of the distributed STAGE kernels — no process group needed:
  * every Distributed* module with world = 1 (all stage kernels run, collectives are no-ops) vs the CPU oracle;
  * the cross-rank combination logic with the ranks emulated IN ONE PROCESS through the C ABI: several row slabs /
    queue shards are pushed through the per-rank stages one after another and combined by the same kernels the
    multi-GPU path uses (the collectives are replaced by torch.cat / sum on the one device).
The real NCCL runs are in tests/test_gpu_dist.py (needs >= 2 GPUs)."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import rel_l2, rel_scalar, ulp_diff
from oracle import ssl_oracle as O
from test_gpu_parity import S, barlow_inputs, check, dev, randn  # noqa: F401  (S is a fixture)

pytestmark = pytest.mark.gpu


def unit(x):
    return (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float32)


# ------------------------------------------------------------------------------------------------ world = 1
@pytest.mark.parametrize("n,d,norm,mode", [(256, 1000, False, "allreduce"), (200, 264, True, "allreduce"),
                                           (512, 4096, False, "allreduce"), (256, 1000, False, "colshard"),
                                           (200, 264, True, "colshard"), (512, 4096, False, "colshard")])
def test_dist_barlow_world1(S, n, d, norm, mode):
    from ssv_b200.dist import DistributedBarlowLoss
    zi, zj = barlow_inputs(n, d)
    ref = O.barlow(zi, zj, norm, 0.005)
    a, b = dev(zi), dev(zj)
    loss = DistributedBarlowLoss(norm, 0.005, mode=mode)(a, b)
    loss.backward()
    check(loss.item(), [a.grad, b.grad], ref[0], ref[1:], f"dist barlow world=1 n={n} d={d}")


@pytest.mark.parametrize("nb,nbank,k,d", [(512, 3000, 3000, 128), (64, 70, 1000, 32), (300, 0, 102, 64)])
def test_dist_swav_world1(S, nb, nbank, k, d):
    from ssv_b200.dist import DistributedSwavLoss
    z1 = unit(randn(0, nb, d))
    z2 = unit(0.6 * z1 + 0.4 * randn(1, nb, d))
    c = unit(randn(2, k, d))
    bank = unit(randn(3, nbank, d)) if nbank else None
    ref = O.swav(z1, z2, c, bank, 0.1, 0.05, 3)
    a, b, p = dev(z1), dev(z2), dev(c)
    loss = DistributedSwavLoss(0.1, 0.05, 3)(a, b, p, dev(bank, False) if nbank else None)
    loss.backward()
    check(loss.item(), [a.grad, b.grad, p.grad], ref[0], ref[1:], f"dist swav world=1 nb={nb} k={k}")


@pytest.mark.parametrize("n,k,d,tau", [(256, 8192, 128, 0.07), (100, 1000, 128, 0.07), (300, 4100, 64, 0.2)])
def test_dist_moco_world1(S, n, k, d, tau):
    from ssv_b200.dist import DistributedMocoLoss
    q, kk = randn(0, n, d), randn(1, n, d)
    mem = unit(randn(2, k, d))
    mem[:7] = 0.0
    ref = O.moco(q, kk, mem, True, tau)
    a, b, m = dev(q), dev(kk), dev(mem, False)
    loss = DistributedMocoLoss(True, tau)(a, b, m)
    loss.backward()
    check(loss.item(), [a.grad, b.grad], ref[0], ref[1:], f"dist moco world=1 n={n} k={k}")


# ------------------------------------------------------------------------------------------------ emulated ranks
@pytest.mark.parametrize("b,k,iters,world", [(4096, 3000, 3, 4), (1024, 5000, 2, 2), (96, 30, 0, 3), (120, 8, 5, 2)])
def test_dist_sinkhorn_emulated_ranks(S, b, k, iters, world):
    """Row slabs of one score matrix pushed through ssvb_sinkhorn_dist_pass as `world` ranks; the all-gather of the
    marginal blocks is a shared [world][k+1] buffer.  Result must equal the oracle on the whole matrix."""
    from ssv_b200 import _cabi as C
    L = C.lib()
    scores = (unit(randn(0, b, 64)) @ unit(randn(1, k, 64)).T).astype(np.float32)
    ref = O.sinkhorn(scores, 0.05, iters)
    s = dev(scores, False)
    bl = b // world
    codes = torch.empty_like(s)
    u_all = torch.empty(world, k + 1, device="cuda")
    alpha = torch.empty(k, device="cuda")
    smax = torch.empty(1, device="cuda")
    nbytes = L.ssvb_sinkhorn_workspace_bytes(bl, k)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    st = C.stream_ptr(s.device)

    def run_pass(phase):
        for r in range(world):
            sl = s[r * bl:(r + 1) * bl]
            C.check(L.ssvb_sinkhorn_dist_pass(phase, C.ptr(sl), bl, b, k, s.stride(0), 0.05,
                                              C.ptr(alpha) if phase else None, C.ptr(smax) if phase else None,
                                              C.ptr(u_all[r]) if phase < 2 else None,
                                              C.ptr(codes[r * bl:(r + 1) * bl]), codes.stride(0),
                                              C.ptr(ws), nbytes, st), "pass")

    run_pass(0)
    C.check(L.ssvb_sinkhorn_dist_alpha(C.ptr(u_all), world, k + 1, k, 1, 0.05, C.ptr(alpha), C.ptr(smax), st), "alpha0")
    assert abs(smax.item() - scores.max()) < 1e-6
    if iters == 0:
        alpha.fill_(1.0)
    for _ in range(1, iters):
        run_pass(1)
        C.check(L.ssvb_sinkhorn_dist_alpha(C.ptr(u_all), world, k + 1, k, 0, 0.05, C.ptr(alpha), C.ptr(smax), st), "alpha")
    run_pass(2)
    got = codes.cpu().numpy()
    assert np.isfinite(got).all()
    assert rel_l2(got, ref) < 1e-4
    np.testing.assert_allclose(got.sum(1), 1.0, rtol=1e-4)


@pytest.mark.parametrize("n,k,d,tau,world", [(64, 8192, 128, 0.07, 4), (50, 1000, 64, 0.2, 2)])
def test_dist_moco_emulated_ranks(S, n, k, d, tau, world):
    """`world` ranks x n queries each against a queue split into `world` shards, all stages through the C ABI in one
    process (all-gather = shared buffers, reduce-scatter = sum of the partial accumulators)."""
    from ssv_b200 import _cabi as C
    L = C.lib()
    ng = n * world
    q, kk = randn(0, ng, d), randn(1, ng, d)
    mem = unit(randn(2, k, d))
    mem[:5] = 0.0
    ref_loss, ref_dq, ref_dk = O.moco(q, kk, mem, True, tau)
    Q, K, M = dev(q, False), dev(kk, False), dev(mem, False)
    kl = k // world
    npad, dpad = L.ssvb_moco_dist_npad(ng), L.ssvb_ntxent_dpad(d)
    st = C.stream_ptr(Q.device)
    qhat_all = torch.empty(npad, dpad, dtype=torch.bfloat16, device="cuda")
    rowstat = torch.empty(world, 3, n, device="cuda")
    for r in range(world):
        C.check(L.ssvb_moco_dist_prep(C.ptr(Q[r * n:]), C.ptr(K[r * n:]), n, d, d, d, 1, world, r, C.ptr(qhat_all),
                                      C.ptr(rowstat[r]), st), "prep")
    nbytes = L.ssvb_moco_dist_workspace_bytes(ng, kl, d)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    part_all = torch.empty(world, 2 * ng + n, device="cuda")
    for r in range(world):
        sh = M[r * kl:(r + 1) * kl]
        C.check(L.ssvb_moco_dist_shard_fwd(C.ptr(qhat_all), ng, C.ptr(sh), None, kl, d, d, tau, C.ptr(rowstat[r]), n,
                                           C.ptr(part_all[r]), C.ptr(ws), nbytes, st), "shard_fwd")
    lse2_all = torch.zeros(npad, device="cuda")
    loss = torch.empty((), device="cuda")
    C.check(L.ssvb_moco_dist_finalize(C.ptr(part_all), world, n, tau, C.ptr(lse2_all), C.ptr(loss), C.ptr(ws), nbytes, st),
            "finalize")
    assert rel_scalar(loss.item(), ref_loss) <= 1e-3
    dacc = torch.zeros(npad, dpad, device="cuda")
    for r in range(world):
        part = torch.empty(npad, dpad, device="cuda")
        sh = M[r * kl:(r + 1) * kl]
        C.check(L.ssvb_moco_dist_shard_bwd(C.ptr(qhat_all), ng, C.ptr(sh), None, kl, d, d, tau, C.ptr(lse2_all),
                                           C.ptr(part), C.ptr(ws), nbytes, st), "shard_bwd")
        dacc[:ng] += part[:ng]
    go = torch.ones((), device="cuda")
    dq, dk = torch.empty(ng, d, device="cuda"), torch.empty(ng, d, device="cuda")
    for r in range(world):
        sl = slice(r * n, (r + 1) * n)
        C.check(L.ssvb_moco_dist_finish(C.ptr(Q[sl]), C.ptr(K[sl]), n, ng, d, d, d, 1, tau, C.ptr(rowstat[r]),
                                        C.ptr(lse2_all[sl]), C.ptr(dacc[sl]), C.ptr(go), C.ptr(dq[sl]), C.ptr(dk[sl]),
                                        d, d, st), "finish")
    assert rel_l2(dq.cpu().numpy(), ref_dq) <= 1e-2
    assert rel_l2(dk.cpu().numpy(), ref_dk) <= 1e-2


def test_sharded_ring_enqueue_bit_exact(S):
    """A ring of 1000 rows split into 4 shards, fed global batches that wrap and (once) exceed the ring: the
    concatenated shards must equal the single-process ring (oracle + the unsharded kernel), pointer included."""
    from ssv_b200 import _cabi as C
    L = C.lib()
    size, d, world = 1000, 32, 4
    rows = size // world
    shards = [torch.zeros(rows, d, device="cuda") for _ in range(world)]
    whole = S.MemoryBank(size, d)
    ref_bank, ref_ptr = np.zeros((size, d), np.float32), 0
    ptr = 0
    st = C.stream_ptr(shards[0].device)
    for step, n in enumerate([300, 300, 300, 300, 1, 1300, 64]):
        batch = randn(10 + step, n, d)
        b = dev(batch, False)
        new_ptrs = []
        for r in range(world):
            np_ = ctypes.c_int64(-1)
            C.check(L.ssvb_ring_enqueue_shard(C.ptr(shards[r]), None, size, r * rows, rows, d, d, C.ptr(b), n, d, ptr, 1,
                                              ctypes.cast(ctypes.pointer(np_), ctypes.c_void_p), st), "enqueue_shard")
            new_ptrs.append(int(np_.value))
        whole.add_batch(b)
        ref_bank, ref_ptr = O.ring_enqueue(ref_bank, ref_ptr, batch, True)
        assert len(set(new_ptrs)) == 1 and new_ptrs[0] == ref_ptr == whole.ptr
        ptr = new_ptrs[0]
        got = torch.cat(shards).cpu().numpy()
        assert np.array_equal(got, whole.bank.cpu().numpy()), f"step {step}: shards differ from the unsharded kernel"
        np.testing.assert_allclose(got, ref_bank, rtol=2.4e-7, atol=0)  # <= 2 ulp of the fp64-rounded oracle (fp32 norm: 1 ulp + 1 ulp for the quotient)
        assert ((got == 0) == (ref_bank == 0)).all()


@pytest.mark.parametrize("n,d,world,norm", [(256, 512, 4, False), (96, 264, 2, True)])
def test_dist_barlow_emulated_ranks(S, n, d, world, norm):
    """Row slabs through the per-rank Barlow stages in one process (stats all-gather = shared buffer, all-reduce of the
    cross-correlation = sum of the partial matrices, slab epilogue per rank)."""
    from ssv_b200 import _cabi as C
    L = C.lib()
    ng = n * world
    zi, zj = barlow_inputs(ng, d)
    ref_loss, ref_di, ref_dj = O.barlow(zi, zj, norm, 0.005)
    Zi, Zj = dev(zi, False), dev(zj, False)
    st = C.stream_ptr(Zi.device)
    sb, wb = L.ssvb_barlow_dist_saved_bytes(n, d), L.ssvb_barlow_dist_workspace_bytes(n, d)
    saved = [torch.empty(sb, dtype=torch.uint8, device="cuda") for _ in range(world)]
    wss = [torch.empty(wb, dtype=torch.uint8, device="cuda") for _ in range(world)]
    stats_all = torch.empty(world, 2, 2, d, device="cuda")
    sl = [slice(r * n, (r + 1) * n) for r in range(world)]
    for r in range(world):
        C.check(L.ssvb_barlow_dist_stats(C.ptr(Zi[sl[r]]), C.ptr(Zj[sl[r]]), n, d, d, d, int(norm), C.ptr(stats_all[r]),
                                         C.ptr(saved[r]), C.ptr(wss[r]), wb, st), "stats")
    csum = torch.zeros(d, d, device="cuda")
    for r in range(world):
        cp = torch.empty(d, d, device="cuda")
        C.check(L.ssvb_barlow_dist_xcorr(C.ptr(Zi[sl[r]]), C.ptr(Zj[sl[r]]), n, d, d, d, int(norm), C.ptr(stats_all), world,
                                         C.ptr(cp), C.ptr(saved[r]), st), "xcorr")
        csum += cp
    dc = torch.empty(d, d, dtype=torch.bfloat16, device="cuda")
    parts = torch.zeros(world, device="cuda")
    rows = d // world
    for r in range(world):
        C.check(L.ssvb_barlow_dist_epilogue(C.ptr(csum[r * rows:]), r * rows, rows, d, 0.005, C.ptr(dc[r * rows:]),
                                            C.ptr(parts[r:]), C.ptr(wss[r]), wb, st), "epilogue")
    assert rel_scalar(parts.sum().item(), ref_loss) <= 1e-3
    colsum = torch.zeros(2, 2, d, device="cuda")
    for r in range(world):
        cl = torch.empty(2, 2, d, device="cuda")
        C.check(L.ssvb_barlow_dist_bwd_gemm(C.ptr(Zi[sl[r]]), C.ptr(Zj[sl[r]]), n, ng, d, d, d, int(norm), C.ptr(dc),
                                            C.ptr(saved[r]), C.ptr(cl), C.ptr(wss[r]), wb, st), "bwd_gemm")
        colsum += cl
    go = torch.ones((), device="cuda")
    di, dj = torch.empty(ng, d, device="cuda"), torch.empty(ng, d, device="cuda")
    for r in range(world):
        C.check(L.ssvb_barlow_dist_bwd_finish(C.ptr(Zi[sl[r]]), C.ptr(Zj[sl[r]]), n, ng, d, d, d, int(norm), C.ptr(colsum),
                                              C.ptr(go), C.ptr(saved[r]), C.ptr(di[sl[r]]), C.ptr(dj[sl[r]]), d, d,
                                              C.ptr(wss[r]), wb, st), "bwd_finish")
    assert rel_l2(di.cpu().numpy(), ref_di) <= 1e-2
    assert rel_l2(dj.cpu().numpy(), ref_dj) <= 1e-2


@pytest.mark.parametrize("n,d,world,norm", [(256, 512, 4, False), (96, 256, 2, True)])
def test_dist_barlow_colshard_emulated_ranks(S, n, d, world, norm):
    """Column-sharded variant with the ranks emulated in one process: the all-gather of the standardised rows is a
    shared [N x D] buffer, the all-to-all of the gradient slabs is a block transpose."""
    from ssv_b200 import _cabi as C
    L = C.lib()
    ng, ncols = n * world, d // world
    zi, zj = barlow_inputs(ng, d)
    ref_loss, ref_di, ref_dj = O.barlow(zi, zj, norm, 0.005)
    Zi, Zj = dev(zi, False), dev(zj, False)
    st = C.stream_ptr(Zi.device)
    sb, wb = L.ssvb_barlow_dist_saved_bytes(n, d), L.ssvb_barlow_dist_workspace_bytes(n, d)
    saved = [torch.empty(sb, dtype=torch.uint8, device="cuda") for _ in range(world)]
    ws = torch.empty(wb, dtype=torch.uint8, device="cuda")
    stats_all = torch.empty(world, 2, 2, d, device="cuda")
    sl = [slice(r * n, (r + 1) * n) for r in range(world)]
    for r in range(world):
        C.check(L.ssvb_barlow_dist_stats(C.ptr(Zi[sl[r]]), C.ptr(Zj[sl[r]]), n, d, d, d, int(norm), C.ptr(stats_all[r]),
                                         C.ptr(saved[r]), C.ptr(ws), wb, st), "stats")
    xt = torch.empty(2, ng, d, dtype=torch.bfloat16, device="cuda")
    for r in range(world):
        C.check(L.ssvb_barlow_dist_standardize(C.ptr(Zi[sl[r]]), C.ptr(Zj[sl[r]]), n, d, d, d, int(norm), C.ptr(stats_all),
                                               world, C.ptr(xt[0, sl[r]]), C.ptr(xt[1, sl[r]]), C.ptr(saved[r]), st),
                "standardize")
    cb = L.ssvb_barlow_cs_workspace_bytes(ng, d, ncols)
    cws = torch.empty(cb, dtype=torch.uint8, device="cuda")
    dc = torch.empty(world, 2, d, ncols, dtype=torch.bfloat16, device="cuda")
    parts = torch.zeros(world, device="cuda")
    for r in range(world):
        C.check(L.ssvb_barlow_cs_fwd(C.ptr(xt[0]), C.ptr(xt[1]), ng, d, r * ncols, ncols, 0.005, C.ptr(dc[r, 0]),
                                     C.ptr(parts[r:]), C.ptr(cws), cb, st), "cs_fwd")
        C.check(L.ssvb_barlow_cs_fwd(C.ptr(xt[1]), C.ptr(xt[0]), ng, d, r * ncols, ncols, 0.005, C.ptr(dc[r, 1]), None,
                                     C.ptr(cws), cb, st), "cs_fwd T")
    assert rel_scalar(parts.sum().item(), ref_loss) <= 1e-3
    go = torch.ones((), device="cuda")
    slabs = torch.empty(world, 2, ng, ncols, device="cuda")  # [owner of the columns][view]
    for r in range(world):
        C.check(L.ssvb_barlow_cs_bwd(C.ptr(xt[0]), C.ptr(xt[1]), C.ptr(dc[r, 0]), ng, d, r * ncols, ncols, C.ptr(saved[r]), n,
                                     1, C.ptr(go), C.ptr(slabs[r, 1]), C.ptr(cws), cb, st), "cs_bwd j")
        C.check(L.ssvb_barlow_cs_bwd(C.ptr(xt[1]), C.ptr(xt[0]), C.ptr(dc[r, 1]), ng, d, r * ncols, ncols, C.ptr(saved[r]), n,
                                     0, C.ptr(go), C.ptr(slabs[r, 0]), C.ptr(cws), cb, st), "cs_bwd i")
    di, dj = torch.empty(ng, d, device="cuda"), torch.empty(ng, d, device="cuda")
    for q in range(world):      # receiver q gets row block q of every owner's slab
        for view, (Z, out) in enumerate(((Zi, di), (Zj, dj))):
            recv = torch.stack([slabs[r, view, sl[q]] for r in range(world)]).contiguous()
            C.check(L.ssvb_barlow_cs_finish(C.ptr(recv), world, n, ncols, C.ptr(Z[sl[q]]), d, int(norm), C.ptr(saved[q]), view,
                                            C.ptr(out[sl[q]]), d, st), "cs_finish")
    assert rel_l2(di.cpu().numpy(), ref_di) <= 1e-2
    assert rel_l2(dj.cpu().numpy(), ref_dj) <= 1e-2


# ------------------------------------------------------------------------------------------------ NT-Xent (headline multi-GPU path)
@pytest.mark.parametrize("world,n,d,norm,tau", [(4, 96, 128, True, 0.5), (2, 256, 64, True, 0.07), (3, 100, 96, False, 1.0),
                                                (8, 64, 128, True, 0.5), (4, 512, 128, True, 0.02)])
def test_dist_ntxent_emulated_ranks(S, world, n, d, norm, tau):
    """The row-sharded NT-Xent stages (ssvb_ntxent_dist_prep / rows_fwd / dist_loss / rows_bwd) with `world` ranks
    emulated in ONE process: every rank's slot is written into one shared gathered matrix (= the all-gather of the
    normalised rows), every rank's [lse | term] block into one statistics buffer (= the second all-gather), and every
    rank's backward reads only those two buffers.  Checked against the fp64 oracle on the rank-order concatenation of
    the inputs (SURVEY.md §8e semantics; reference math utils/losses.py:15-46): loss, and each rank's gradient rows."""
    from ssv_b200 import _cabi as C
    from ssv_b200.dist import CudaStages
    stg = CudaStages()
    zi, zj = randn(0, world * n, d), randn(1, world * n, d)
    if 0.05 < tau < 0.1:  # peaky softmax: correlated positives (at tau = 0.02 that would make the loss ~1e-9)
        zj = (0.8 * zi + 0.6 * zj).astype(np.float32)
    ref_loss, ref_di, ref_dj = O.ntxent(zi, zj, norm, tau)
    Zi, Zj = dev(zi, False), dev(zj, False)
    mpad, dpad = stg.mpad(n * world), stg.dpad(d)
    zhat_all = torch.zeros(mpad, dpad, dtype=torch.bfloat16, device="cuda")
    stat_all = torch.empty(world, 2, 2 * n, dtype=torch.float32, device="cuda")
    inv = [torch.empty(2 * n, device="cuda") for _ in range(world)]
    pos = [torch.empty(2 * n, device="cuda") for _ in range(world)]
    sums = [torch.zeros((), device="cuda") for _ in range(world)]
    sl = [slice(r * n, (r + 1) * n) for r in range(world)]
    for r in range(world):
        stg.prep(Zi[sl[r]], Zj[sl[r]], int(norm), tau, world, r, zhat_all, inv[r], pos[r])
    for r in range(world):
        stg.rows_fwd(zhat_all, world, r, n, d, int(norm), tau, pos[r], stat_all[r], sums[r])
    loss = torch.empty((), device="cuda")
    stg.dist_loss(stat_all, world, n, loss)
    m = 2 * n * world
    assert rel_scalar(loss.item(), ref_loss) <= 1e-3
    assert rel_scalar(sum(s.item() for s in sums) / m, ref_loss) <= 1e-3   # the NCCL transport's world = 1 reduction
    go = torch.full((), 1.5, device="cuda")
    for r in range(world):
        dzi, dzj = torch.empty(n, d, device="cuda"), torch.empty(n, d, device="cuda")
        stg.rows_bwd(Zi[sl[r]], Zj[sl[r]], int(norm), tau, world, r, zhat_all, stat_all, inv[r], go, dzi, dzj)
        assert rel_l2(dzi.cpu().numpy(), 1.5 * ref_di[sl[r]]) <= 1e-2, f"rank {r} dzi"
        assert rel_l2(dzj.cpu().numpy(), 1.5 * ref_dj[sl[r]]) <= 1e-2, f"rank {r} dzj"


@pytest.mark.parametrize("world,n,d,norm,tau", [(4, 96, 128, True, 0.5), (8, 64, 128, True, 0.5), (2, 300, 64, True, 0.07),
                                                (3, 100, 96, False, 1.0)])
def test_dist_ntxent_p2p_emulated_ranks(S, world, n, d, norm, tau):
    """The NVLink peer-memory transport (ssvb_ntxent_p2p_*: pushes into every rank's arena + generation flags) with
    the ranks emulated on ONE GPU: `world` arenas in the same device memory stand in for the peer-mapped arenas, the
    kernels are the ones the multi-GPU path launches (unicast stores).  Three generations with different inputs
    exercise both buffer parities and the monotonic flags; the loss must be bit-identical on every rank."""
    from ssv_b200 import _cabi as C
    from ssv_b200.dist import CudaStages
    stg = CudaStages()
    nbytes = C.lib().ssvb_ntxent_p2p_arena_bytes(world, n, d)
    arenas = [torch.zeros(nbytes, dtype=torch.uint8, device="cuda") for _ in range(world)]
    peers = torch.tensor([a.data_ptr() for a in arenas], dtype=torch.int64, device="cuda")

    class Arena:
        def __init__(self, r):
            self.local_ptr, self.peers_dev, self.multicast_ptr = arenas[r].data_ptr(), peers.data_ptr(), 0
    ar = [Arena(r) for r in range(world)]
    mpad, dpad = stg.mpad(n * world), stg.dpad(d)
    sl = [slice(r * n, (r + 1) * n) for r in range(world)]
    for gen in (1, 2, 3):
        zi, zj = randn(10 * gen, world * n, d), randn(10 * gen + 1, world * n, d)
        if tau < 0.1:
            zj = (0.8 * zi + 0.6 * zj).astype(np.float32)
        ref_loss, ref_di, ref_dj = O.ntxent(zi, zj, norm, tau)
        Zi, Zj = dev(zi, False), dev(zj, False)
        inv = [torch.empty(2 * n, device="cuda") for _ in range(world)]
        pos = [torch.empty(2 * n, device="cuda") for _ in range(world)]
        zh = [torch.empty(mpad, dpad, dtype=torch.bfloat16, device="cuda") for _ in range(world)]
        cs = [torch.empty(mpad, device="cuda") for _ in range(world)]
        losses = [torch.empty((), device="cuda") for _ in range(world)]
        tmp = torch.empty((), device="cuda")
        for r in range(world):
            stg.p2p_prep_push(Zi[sl[r]], Zj[sl[r]], int(norm), tau, world, r, ar[r], gen, inv[r], pos[r])
        for r in range(world):
            stg.p2p_wait_copy(ar[r], world, r, n, d, gen, zh[r])
        for r in range(1, world):
            assert torch.equal(zh[0][:2 * n * world], zh[r][:2 * n * world]), "every rank must gather the same rows"
        for r in range(world):
            stg.p2p_rows_fwd(zh[r], world, r, n, d, int(norm), tau, pos[r], ar[r], gen, tmp)
        for r in range(world):
            stg.p2p_stat_loss(ar[r], world, r, n, d, int(norm), tau, gen, cs[r], losses[r])
        torch.cuda.synchronize()
        assert all(torch.equal(losses[0], l) for l in losses), "the global loss must be bit-identical on every rank"
        assert rel_scalar(losses[0].item(), ref_loss) <= 1e-3, f"gen {gen}"
        go = torch.full((), 0.5, device="cuda")
        for r in range(world):
            dzi, dzj = torch.empty(n, d, device="cuda"), torch.empty(n, d, device="cuda")
            stg.p2p_rows_bwd(Zi[sl[r]], Zj[sl[r]], int(norm), tau, world, r, zh[r], cs[r], inv[r], go, dzi, dzj)
            assert rel_l2(dzi.cpu().numpy(), 0.5 * ref_di[sl[r]]) <= 1e-2, f"gen {gen} rank {r} dzi"
            assert rel_l2(dzj.cpu().numpy(), 0.5 * ref_dj[sl[r]]) <= 1e-2, f"gen {gen} rank {r} dzj"


# ------------------------------------------------------------------------------------------------ ReLIC (KL over the global batch)
@pytest.mark.parametrize("n,d,tau,alpha", [(512, 128, 1.0, 0.5), (300, 64, 0.5, 2.0)])
def test_dist_relic_world1(S, n, d, tau, alpha):
    from ssv_b200.dist import DistributedRelicLoss
    zi, zj, zo = randn(0, n, d), randn(1, n, d), randn(2, n, d)
    ref = O.relic(zi, zj, zo, True, tau, alpha)
    a, b, c = dev(zi), dev(zj), dev(zo)
    loss = DistributedRelicLoss(True, tau, alpha)(a, b, c)
    loss.backward()
    check(loss.item(), [a.grad, b.grad, c.grad], ref[0], ref[1:], f"dist relic world=1 n={n}")


@pytest.mark.parametrize("world,n,d,tau,alpha", [(4, 96, 128, 1.0, 0.5), (3, 100, 64, 0.5, 1.5)])
def test_dist_relic_kl_emulated_ranks(S, world, n, d, tau, alpha):
    """The KL stages of the distributed ReLIC loss with the ranks emulated in one process: per-rank dots, the all-gather
    = one shared [world][2][n] buffer, per-rank reduce + backward.  KL value and gradients = oracle RelicLoss minus
    oracle NT-Xent (the contrastive part) on the concatenation (reference utils/losses.py:196-200)."""
    from ssv_b200.dist import RelicKlCudaStages
    stg = RelicKlCudaStages()
    zi, zj, zo = randn(0, world * n, d), randn(1, world * n, d), randn(2, world * n, d)
    full = O.relic(zi, zj, zo, True, tau, alpha)
    con = O.ntxent(zi, zj, True, tau)
    ref_kl, ref_di, ref_dj, ref_do = full[0] - con[0], full[1] - con[1], full[2] - con[2], full[3]
    Zi, Zj, Zo = dev(zi, False), dev(zj, False), dev(zo, False)
    sl = [slice(r * n, (r + 1) * n) for r in range(world)]
    ab_all = torch.empty(world, 2, n, device="cuda")
    saved = [stg.alloc_saved(n, Zi.device) for _ in range(world)]
    for r in range(world):
        stg.dots(Zi[sl[r]], Zj[sl[r]], Zo[sl[r]], 1, tau, saved[r], ab_all[r])
    kls = [torch.empty((), device="cuda") for _ in range(world)]
    for r in range(world):
        stg.reduce(ab_all, world, n, alpha, saved[r], kls[r])
    assert all(torch.equal(kls[0], k) for k in kls)
    assert rel_scalar(kls[0].item(), ref_kl) <= 1e-3
    go = torch.ones((), device="cuda")
    for r in range(world):
        dzi, dzj, dzo = torch.zeros(n, d, device="cuda"), torch.zeros(n, d, device="cuda"), torch.empty(n, d, device="cuda")
        stg.bwd(Zi[sl[r]], Zj[sl[r]], Zo[sl[r]], 1, tau, alpha, go, saved[r], dzi, dzj, dzo)
        assert rel_l2(dzi.cpu().numpy(), ref_di[sl[r]]) <= 1e-2
        assert rel_l2(dzj.cpu().numpy(), ref_dj[sl[r]]) <= 1e-2
        assert rel_l2(dzo.cpu().numpy(), ref_do[sl[r]]) <= 1e-2


def test_dist_dino_center_world1(S):
    """distributed_update_teacher_center without a process group == the single-GPU kernel == the oracle
    (reference models/dino.py:136-141)."""
    from ssv_b200.dist import distributed_update_teacher_center
    t = randn(0, 64, 2, 1024)
    c0 = 0.1 * randn(1, 1024)
    ref_first = O.dino_center_update(None, t.reshape(-1, 1024), 0.9)
    ref = O.dino_center_update(c0, t.reshape(-1, 1024), 0.9)
    got_first = distributed_update_teacher_center(None, dev(t.reshape(-1, 1024), False), 0.9)
    got = distributed_update_teacher_center(dev(c0, False), dev(t.reshape(-1, 1024), False), 0.9)
    np.testing.assert_allclose(got_first.cpu().numpy(), ref_first, rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=2e-5, atol=1e-6)
