"""world_size-2 gloo test of the multi-GPU orchestration (ssv_b200/dist.py) on CPU.

The three compute stages are replaced by a numpy emulation with the same contract as the C ABI
(ssvb_ntxent_dist_prep / rows_fwd / rows_bwd, rank-major gathered layout); what is under test is the host logic:
slot layout, the two all-gathers, the loss all-reduce and scaling, and that every rank ends up with exactly the
gradient rows of its own inputs — all against the single-process oracle on the concatenated global batch."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class EmulatedStages:
    """numpy stand-in for the CUDA stages (test-only)."""

    def dpad(self, d):
        return d

    def mpad(self, n_global):
        return 2 * n_global

    def prep(self, zi, zj, normalize, temperature, world, rank, zhat_all, inv_local, pos_local):
        n = zi.shape[0]
        z = torch.cat([zi, zj]).double()
        den = z.norm(dim=1, keepdim=True).clamp_min(1e-12) if normalize else torch.ones(2 * n, 1, dtype=torch.float64)
        zh = z / den
        zhat_all[rank * 2 * n:(rank + 1) * 2 * n] = zh.to(zhat_all.dtype)
        inv_local.copy_((1.0 / den[:, 0]).float())
        pos = (zh[:n] * zh[n:]).sum(1)
        pos_local.copy_(torch.cat([pos, pos]).float())

    def rows_fwd(self, zhat_all, world, rank, n_local, d, normalize, temperature, pos_local, stat_local, loss_sum):
        lr = 2 * n_local
        z = zhat_all.double()
        s = z[rank * lr:(rank + 1) * lr] @ z.t() / temperature
        idx = torch.arange(lr)
        s[idx, rank * lr + idx] = float("-inf")
        lse = torch.logsumexp(s, 1)
        term = lse - pos_local.double() / temperature
        stat_local[0].copy_(lse.float())     # stat_local: [2][2L] = LSE, then per-row loss terms
        stat_local[1].copy_(term.float())
        loss_sum.copy_(term.sum().float())

    def rows_bwd(self, zi, zj, normalize, temperature, world, rank, zhat_all, stat_all, inv_local, grad_out, dzi, dzj):
        n = zi.shape[0]
        lr = 2 * n
        m = zhat_all.shape[0]
        z = zhat_all.double()
        zl = z[rank * lr:(rank + 1) * lr]
        s = zl @ z.t() / temperature
        lse_all = stat_all.double()[:, 0, :].reshape(-1)   # gathered layout [world][2][2L]
        w = torch.exp(s - lse_all[rank * lr:(rank + 1) * lr, None]) + torch.exp(s - lse_all[None, :])
        idx = torch.arange(lr)
        w[idx, rank * lr + idx] = 0.0
        partner = rank * lr + (idx + n) % lr
        g = (w @ z - 2.0 * z[partner]) / (m * temperature) * grad_out.double()
        if normalize:
            g = (g - (g * zl).sum(1, keepdim=True) * zl) * inv_local.double()[:, None]
        dzi.copy_(g[:n].float())
        dzj.copy_(g[n:].float())


def _worker(rank, world, port, n_local, d, normalize, tau, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssv_b200.dist import DistributedSimclrLoss
    g = torch.Generator().manual_seed(100 + rank)
    zi = torch.randn(n_local, d, generator=g, requires_grad=True)
    zj = torch.randn(n_local, d, generator=g, requires_grad=True)
    loss = DistributedSimclrLoss(normalize, tau, stages=EmulatedStages())(zi, zj)
    (2.0 * loss).backward()  # also checks that grad_out reaches the stages
    out[rank] = (loss.item(), zi.grad.numpy().copy(), zj.grad.numpy().copy(), zi.detach().numpy().copy(),
                 zj.detach().numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("normalize,tau", [(True, 0.5), (False, 1.0)])
def test_distributed_ntxent_matches_global_oracle(normalize, tau):
    from oracle import ssl_oracle as O
    world, n_local, d = 2, 24, 16
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n_local, d, normalize, tau, out), nprocs=world, join=True)
    zi = np.concatenate([out[r][3] for r in range(world)]) * (1.0 if normalize else 0.3)
    zj = np.concatenate([out[r][4] for r in range(world)]) * (1.0 if normalize else 0.3)
    if not normalize:  # the workers used unscaled inputs; recompute the oracle on exactly what they saw
        zi, zj = zi / 0.3, zj / 0.3
    ref_loss, ref_dzi, ref_dzj = O.ntxent(zi, zj, normalize, tau)
    for r in range(world):
        loss, gi, gj, _, _ = out[r]
        assert abs(loss - ref_loss) / abs(ref_loss) < 1e-4  # bf16 staging of zhat in the emulated gather
        sl = slice(r * n_local, (r + 1) * n_local)
        assert np.linalg.norm(gi - 2 * ref_dzi[sl]) / np.linalg.norm(2 * ref_dzi[sl]) < 2e-2
        assert np.linalg.norm(gj - 2 * ref_dzj[sl]) / np.linalg.norm(2 * ref_dzj[sl]) < 2e-2
    assert out[0][0] == out[1][0], "every rank must report the same global loss"


# ======================================================================================================= Barlow Twins
class EmulatedBarlowStages:
    """torch-fp64 stand-in for the ssvb_barlow_dist_* stages (test-only): same contracts, same buffers."""

    def alloc_saved(self, n, d, dev):
        return torch.zeros(2 * n * d + 4 * d + 2 * n, dtype=torch.float64)

    @staticmethod
    def _views(saved, n, d):
        o = 0
        out = []
        for cnt, shape in ((n * d, (n, d)), (n * d, (n, d)), (d, (d,)), (d, (d,)), (d, (d,)), (d, (d,)), (n, (n,)), (n, (n,))):
            out.append(saved[o:o + cnt].view(*shape))
            o += cnt
        return out  # xi~, xj~, mean_i, rstd_i, mean_j, rstd_j, inv_i, inv_j

    def stats(self, zi, zj, normalize, stats_local, saved):
        n, d = zi.shape
        v = self._views(saved, n, d)
        for k, z in enumerate((zi, zj)):
            x = z.double()
            inv = 1.0 / x.norm(dim=1).clamp_min(1e-12) if normalize else torch.ones(n, dtype=torch.float64)
            v[6 + k].copy_(inv)
            x = x * inv[:, None]
            mu = x.mean(0)
            stats_local[k, 0].copy_(mu.float())
            stats_local[k, 1].copy_(((x - mu) ** 2).sum(0).float())

    def xcorr(self, zi, zj, normalize, stats_all, world, c_partial, saved):
        n, d = zi.shape
        v = self._views(saved, n, d)
        ng = n * world
        xt = []
        for k, z in enumerate((zi, zj)):
            mu_r = stats_all[:, k, 0].double()
            m2_r = stats_all[:, k, 1].double()
            mu = mu_r.mean(0)
            m2 = m2_r.sum(0) + n * ((mu_r - mu) ** 2).sum(0)
            rstd = 1.0 / torch.sqrt(m2 / (ng - 1))
            v[2 + 2 * k].copy_(mu)
            v[3 + 2 * k].copy_(rstd)
            x = z.double() * v[6 + k][:, None]
            v[k].copy_((x - mu) * rstd)
            xt.append(v[k])
        c_partial.copy_((xt[0].t() @ xt[1] / ng).float())

    def epilogue(self, c_rows, row0, lmbda, dc_rows, loss_partial, n):
        rows, d = c_rows.shape
        c = c_rows.double()
        eye = torch.zeros(rows, d, dtype=torch.bool)
        eye[torch.arange(rows), row0 + torch.arange(rows)] = True
        t = torch.where(eye, c - 1.0, c)
        w = torch.where(eye, torch.ones_like(c), torch.full_like(c, lmbda))
        loss_partial.copy_((w * t * t).sum().float().view(1))
        dc_rows.copy_((2.0 * w * t).to(dc_rows.dtype))

    def bwd_gemm(self, zi, zj, n_global, normalize, dc, saved, colsum):
        n, d = zi.shape
        v = self._views(saved, n, d)
        g = dc.double()
        self.dti = v[1] @ g.t() / n_global
        self.dtj = v[0] @ g / n_global
        colsum[0, 0].copy_(self.dti.sum(0).float())
        colsum[0, 1].copy_((self.dti * v[0]).sum(0).float())
        colsum[1, 0].copy_(self.dtj.sum(0).float())
        colsum[1, 1].copy_((self.dtj * v[1]).sum(0).float())

    def bwd_finish(self, zi, zj, n_global, normalize, colsum, grad_out, saved, dzi, dzj):
        n, d = zi.shape
        v = self._views(saved, n, d)
        for k, (z, dt, out) in enumerate(((zi, self.dti, dzi), (zj, self.dtj, dzj))):
            m1 = colsum[k, 0].double() / n_global
            m2 = colsum[k, 1].double() / (n_global - 1)
            g = (dt - m1 - v[k] * m2) * v[3 + 2 * k] * grad_out.double()
            if normalize:
                xh = z.double() * v[6 + k][:, None]
                g = (g - (g * xh).sum(1, keepdim=True) * xh) * v[6 + k][:, None]
            out.copy_(g.float())


    # ---- column-sharded variant
    def standardize(self, zi, zj, normalize, stats_all, world, xi_slot, xj_slot, saved):
        n, d = zi.shape
        v = self._views(saved, n, d)
        ng = n * world
        for k, (z, slot) in enumerate(((zi, xi_slot), (zj, xj_slot))):
            mu_r, m2_r = stats_all[:, k, 0].double(), stats_all[:, k, 1].double()
            mu = mu_r.mean(0)
            m2 = m2_r.sum(0) + n * ((mu_r - mu) ** 2).sum(0)
            rstd = 1.0 / torch.sqrt(m2 / (ng - 1))
            v[2 + 2 * k].copy_(mu)
            v[3 + 2 * k].copy_(rstd)
            slot.copy_((((z.double() * v[6 + k][:, None]) - mu) * rstd).to(slot.dtype))

    def cs_fwd(self, xa_all, xb_all, col0, ncols, lmbda, dc_slab, loss_partial):
        ng, d = xa_all.shape
        c = xa_all.double().t() @ xb_all.double()[:, col0:col0 + ncols] / ng
        eye = torch.zeros(d, ncols, dtype=torch.bool)
        eye[col0 + torch.arange(ncols), torch.arange(ncols)] = True
        t = torch.where(eye, c - 1.0, c)
        w = torch.where(eye, torch.ones_like(c), torch.full_like(c, lmbda))
        dc_slab.copy_((2.0 * w * t).to(dc_slab.dtype))
        if loss_partial is not None:
            loss_partial.copy_((w * t * t).sum().float().view(1))

    def cs_bwd(self, xa_all, xb_all, dc_slab, col0, ncols, saved, n_local, view_b, grad_out, dxb_slab):
        ng, d = xa_all.shape
        v = self._views(saved, n_local, d)
        dt = xa_all.double() @ dc_slab.double() / ng
        xb = xb_all.double()[:, col0:col0 + ncols]
        m1 = dt.mean(0)
        m2 = (dt * xb).sum(0) / (ng - 1)
        dxb_slab.copy_(((dt - m1 - xb * m2) * v[3 + 2 * view_b][col0:col0 + ncols] * grad_out.double()).float())

    def cs_finish(self, recv, world, n_local, ncols, x, normalize, saved, view, dx):
        d = world * ncols
        v = self._views(saved, n_local, d)
        g = recv.double().permute(1, 0, 2).reshape(n_local, d)
        if normalize:
            xh = x.double() * v[6 + view][:, None]
            g = (g - (g * xh).sum(1, keepdim=True) * xh) * v[6 + view][:, None]
        dx.copy_(g.float())


def _barlow_worker(rank, world, port, n_local, d, normalize, lmbda, out, mode="allreduce"):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssv_b200.dist import DistributedBarlowLoss
    g = torch.Generator().manual_seed(200 + rank)
    zi = (torch.randn(n_local, d, generator=g) * 1.5 + 0.3).requires_grad_(True)
    zj = (zi.detach() * 0.7 + 0.5 * torch.randn(n_local, d, generator=g)).requires_grad_(True)
    loss = DistributedBarlowLoss(normalize, lmbda, stages=EmulatedBarlowStages(), mode=mode)(zi, zj)
    (3.0 * loss).backward()
    out[rank] = (loss.item(), zi.grad.numpy().copy(), zj.grad.numpy().copy(), zi.detach().numpy().copy(),
                 zj.detach().numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


# d=33: not divisible by the world size -> plain all-reduce of C instead of reduce-scatter + all-gather
@pytest.mark.parametrize("normalize,d,mode", [(False, 32, "allreduce"), (True, 32, "allreduce"), (False, 33, "allreduce"),
                                              (False, 32, "colshard"), (True, 32, "colshard")])
def test_distributed_barlow_matches_global_oracle(normalize, d, mode):
    from oracle import ssl_oracle as O
    world, n_local, lmbda = 2, 20, 0.005
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_barlow_worker, args=(world, port, n_local, d, normalize, lmbda, out, mode), nprocs=world, join=True)
    zi = np.concatenate([out[r][3] for r in range(world)])
    zj = np.concatenate([out[r][4] for r in range(world)])
    ref_loss, ref_dzi, ref_dzj = O.barlow(zi, zj, normalize, lmbda)
    for r in range(world):
        loss, gi, gj, _, _ = out[r]
        assert abs(loss - ref_loss) / abs(ref_loss) < 1e-3
        sl = slice(r * n_local, (r + 1) * n_local)
        assert np.linalg.norm(gi - 3 * ref_dzi[sl]) / np.linalg.norm(3 * ref_dzi[sl]) < 2e-2  # dC staged in bf16
        assert np.linalg.norm(gj - 3 * ref_dzj[sl]) / np.linalg.norm(3 * ref_dzj[sl]) < 2e-2
    assert out[0][0] == out[1][0], "every rank must report the same global loss"


# ======================================================================================================= SwAV / Sinkhorn
class EmulatedSwavStages:
    """torch-fp64 stand-in for the distributed SwAV / Sinkhorn C-ABI stages (test-only)."""

    def kpad(self, k):
        return (k + 3) // 4 * 4

    def alloc_saved(self, nb, nbank, k, d, dev):
        return torch.zeros(1)

    def sk_pass(self, phase, scores, b_global, k, eps, alpha, smax, u_local, codes):
        s = scores[:, :k].double()
        if phase == 0:
            m = s.max()
            u_local[:k].copy_(torch.exp((s - m) / eps).sum(0).float())
            u_local[k] = m.float()
            return
        e = torch.exp((s - smax.double()) / eps)
        v = e @ alpha.double()
        if phase == 1:
            u_local[:k].copy_((e / (b_global * v[:, None])).sum(0).float())
            u_local[k] = 0.0
        else:
            codes[:, :k].copy_((e * alpha.double() / v[:, None]).float())

    def sk_alpha(self, u_all_view, world, rank_stride, k, phase0, eps, alpha, smax):
        blocks = u_all_view.as_strided((world, k + 1), (rank_stride, 1)).double()
        if phase0:
            m = blocks[:, k].max()
            u = (blocks[:, :k] * torch.exp((blocks[:, k] - m) / eps)[:, None]).sum(0)
            smax.copy_(m.float().view(1))
        else:
            u = blocks[:, :k].sum(0)
        alpha.copy_(((1.0 / k) / u).float())

    def scores(self, z1, z2, bank, protos, scores, saved):
        rows = [torch.cat([z, bank]) if bank is not None else z for z in (z1, z2)]
        self.z = torch.cat(rows).double()
        self.c = protos.double()
        scores[:, :protos.shape[0]].copy_((self.z @ self.c.t()).float())

    def ce(self, scores, codes, nb, nbank, bp_global, k, d, temperature, loss_local, saved):
        bp = nb + nbank
        s = scores[:, :k].double() / temperature
        q = codes[:, :k].double()
        p = torch.log_softmax(s, 1)
        q1, q2, p1, p2 = q[:bp], q[bp:], p[:bp], p[bp:]
        loss_local.copy_((-0.5 * ((q1 * p2).sum() + (q2 * p1).sum()) / bp_global).float())
        coef = 0.5 / (bp_global * temperature)
        ds1 = -(q2 - torch.exp(p1) * q2.sum(1, keepdim=True)) * coef
        ds2 = -(q1 - torch.exp(p2) * q1.sum(1, keepdim=True)) * coef
        self.ds = torch.cat([ds1, ds2])
        self.bp, self.nb = bp, nb

    def bwd(self, z1, z2, bank, protos, temperature, grad_out, saved, dz1, dz2, dproto):
        go = grad_out.double()
        dz = self.ds @ self.c * go
        dz1.copy_(dz[:self.nb].float())
        dz2.copy_(dz[self.bp:self.bp + self.nb].float())
        dproto.copy_((self.ds.t() @ self.z * go).float())


def _swav_worker(rank, world, port, nb, nbank, k, d, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssv_b200.dist import DistributedSwavLoss
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(300 + rank)
    z1 = F.normalize(torch.randn(nb, d, generator=g)).requires_grad_(True)
    z2 = F.normalize(torch.randn(nb, d, generator=g)).requires_grad_(True)
    bank = F.normalize(torch.randn(nbank, d, generator=g)) if nbank else None
    gp = torch.Generator().manual_seed(7)  # prototypes are replicated: same on every rank
    pc = F.normalize(torch.randn(k, d, generator=gp)).requires_grad_(True)
    fn = DistributedSwavLoss(0.1, 0.05, 3, stages=EmulatedSwavStages())
    loss = fn(z1, z2, pc, bank)
    (2.0 * loss).backward()
    sc = (z1.detach() @ pc.detach().t())
    codes = fn.compute_codes_sinkhorn(sc)
    out[rank] = dict(loss=loss.item(), dz1=z1.grad.numpy().copy(), dz2=z2.grad.numpy().copy(), dpc=pc.grad.numpy().copy(),
                     z1=z1.detach().numpy().copy(), z2=z2.detach().numpy().copy(),
                     bank=None if bank is None else bank.numpy().copy(), pc=pc.detach().numpy().copy(),
                     sc=sc.numpy().copy(), codes=codes.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nbank", [0, 12])
def test_distributed_swav_matches_global_oracle(nbank):
    from oracle import ssl_oracle as O
    world, nb, k, d = 2, 16, 10, 8
    mgr = mp.Manager()
    out = mgr.dict()
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_swav_worker, args=(world, port, nb, nbank, k, d, out), nprocs=world, join=True)
    # the single-process reference sees: live rows of all ranks, then bank rows of all ranks (row order within a view is
    # irrelevant to the loss as long as both views use the same order)
    z1 = np.concatenate([out[r]["z1"] for r in range(world)])
    z2 = np.concatenate([out[r]["z2"] for r in range(world)])
    bank = np.concatenate([out[r]["bank"] for r in range(world)]) if nbank else None
    ref_loss, ref_dz1, ref_dz2, ref_dc = O.swav(z1, z2, out[0]["pc"], bank, 0.1, 0.05, 3)
    for r in range(world):
        o = out[r]
        assert abs(o["loss"] - ref_loss) / abs(ref_loss) < 1e-4
        sl = slice(r * nb, (r + 1) * nb)
        assert np.linalg.norm(o["dz1"] - 2 * ref_dz1[sl]) / np.linalg.norm(2 * ref_dz1[sl]) < 1e-3
        assert np.linalg.norm(o["dz2"] - 2 * ref_dz2[sl]) / np.linalg.norm(2 * ref_dz2[sl]) < 1e-3
    # each rank returns its LOCAL contribution to the (replicated) prototype gradient, consistent with the
    # row gradients: the SUM over the ranks - what DDP / an all-reduce of the replicated parameters computes - is the
    # single-process gradient, and encoder rows and prototypes carry the same scale
    dpc_sum = sum(out[r]["dpc"].astype(np.float64) for r in range(world))
    assert np.linalg.norm(dpc_sum - 2 * ref_dc) / np.linalg.norm(2 * ref_dc) < 1e-3
    assert not np.allclose(out[0]["dpc"], out[1]["dpc"]), "per-rank contributions differ (they are not pre-reduced)"
    assert out[0]["loss"] == out[1]["loss"]
    # stand-alone distributed Sinkhorn: codes of the row-sharded score matrix == rows of the global codes
    ref_codes = O.sinkhorn(np.concatenate([out[r]["sc"] for r in range(world)]), 0.05, 3)
    for r in range(world):
        assert np.allclose(out[r]["codes"], ref_codes[r * nb:(r + 1) * nb], rtol=1e-4, atol=1e-7)


# ======================================================================================================= MoCo (sharded queue)
class EmulatedMocoStages:
    """torch-fp64 stand-in for the ssvb_moco_dist_* stages (test-only): log2-domain (max, sum) shard partials."""
    LOG2E = 1.4426950408889634

    def dpad(self, d):
        return d

    def npad(self, n_global):
        return n_global

    def prep(self, q, k, normalize, world, rank, qhat_all, rowstat):
        n = q.shape[0]
        qd, kd = q.double(), k.double()
        iq = 1.0 / qd.norm(dim=1).clamp_min(1e-12) if normalize else torch.ones(n, dtype=torch.float64)
        ik = 1.0 / kd.norm(dim=1).clamp_min(1e-12) if normalize else torch.ones(n, dtype=torch.float64)
        qh, kh = qd * iq[:, None], kd * ik[:, None]
        qhat_all[rank * n:(rank + 1) * n] = qh.to(qhat_all.dtype)
        rowstat[0].copy_(iq.float())
        rowstat[1].copy_(ik.float())
        rowstat[2].copy_((qh * kh).sum(1).float())

    def shard_fwd(self, qhat_all, n_global, shard, shadow, d, temperature, rowstat, n_local, part_local):
        c = self.LOG2E / temperature
        t = (qhat_all[:n_global].double() @ shard.double().t()) * c
        m = t.max(1).values
        part_local[:n_global].copy_(m.float())
        part_local[n_global:2 * n_global].copy_(torch.exp2(t - m[:, None]).sum(1).float())
        part_local[2 * n_global:].copy_(rowstat[2])

    def finalize(self, part_all, world, n_local, temperature, lse2_all, loss, k_local, d):
        ng = world * n_local
        c = self.LOG2E / temperature
        pa = part_all.double()
        p2 = torch.cat([pa[w, 2 * ng:] for w in range(world)]) * c
        m = torch.maximum(pa[:, :ng].max(0).values, p2)
        l = torch.exp2(p2 - m) + (pa[:, ng:2 * ng] * torch.exp2(pa[:, :ng] - m)).sum(0)
        lse2 = m + torch.log2(l)
        lse2_all[:ng].copy_(lse2.float())
        loss.copy_((((lse2 - p2) / self.LOG2E).sum() / ng).float())

    def shard_bwd(self, qhat_all, n_global, shard, shadow, d, temperature, lse2_all, dacc_partial):
        c = self.LOG2E / temperature
        t = (qhat_all[:n_global].double() @ shard.double().t()) * c
        p = torch.exp2(t - lse2_all[:n_global].double()[:, None])
        dacc_partial[:n_global].copy_((p @ shard.double()).float())

    def finish(self, q, k, n_global, normalize, temperature, rowstat, lse2_local, dacc_local, grad_out, dq, dk):
        c = self.LOG2E / temperature
        iq, ik, pos = rowstat[0].double(), rowstat[1].double(), rowstat[2].double()
        qh, kh = q.double() * iq[:, None], k.double() * ik[:, None]
        p0m1 = (torch.exp2(pos * c - lse2_local.double()) - 1.0)[:, None]
        scale = grad_out.double() / (n_global * temperature)
        gq = (p0m1 * kh + dacc_local.double()) * scale
        gk = p0m1 * qh * scale
        if normalize:
            gq = (gq - (gq * qh).sum(1, keepdim=True) * qh) * iq[:, None]
            gk = (gk - (gk * kh).sum(1, keepdim=True) * kh) * ik[:, None]
        dq.copy_(gq.float())
        dk.copy_(gk.float())


def _moco_worker(rank, world, port, n_local, k_total, d, tau, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssv_b200.dist import DistributedMocoLoss
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(400 + rank)
    q = torch.randn(n_local, d, generator=g, requires_grad=True)
    k = torch.randn(n_local, d, generator=g, requires_grad=True)
    gq = torch.Generator().manual_seed(9)
    queue = F.normalize(torch.randn(k_total, d, generator=gq))
    queue[3] = 0.0  # a never-written (zero) row, as in a fresh MemoryBank
    kl = k_total // world
    shard = queue[rank * kl:(rank + 1) * kl].contiguous()
    loss = DistributedMocoLoss(True, tau, stages=EmulatedMocoStages())(q, k, shard)
    (0.5 * loss).backward()
    out[rank] = dict(loss=loss.item(), dq=q.grad.numpy().copy(), dk=k.grad.numpy().copy(), q=q.detach().numpy().copy(),
                     k=k.detach().numpy().copy(), queue=queue.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("tau", [0.07, 1.0])
def test_distributed_moco_sharded_queue_matches_global_oracle(tau):
    from oracle import ssl_oracle as O
    world, n_local, k_total, d = 2, 12, 40, 16
    mgr = mp.Manager()
    out = mgr.dict()
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_moco_worker, args=(world, port, n_local, k_total, d, tau, out), nprocs=world, join=True)
    q = np.concatenate([out[r]["q"] for r in range(world)])
    k = np.concatenate([out[r]["k"] for r in range(world)])
    ref_loss, ref_dq, ref_dk = O.moco(q, k, out[0]["queue"], True, tau)
    for r in range(world):
        o = out[r]
        assert abs(o["loss"] - ref_loss) / abs(ref_loss) < 1e-3  # queries travel in bf16
        sl = slice(r * n_local, (r + 1) * n_local)
        assert np.linalg.norm(o["dq"] - 0.5 * ref_dq[sl]) / np.linalg.norm(0.5 * ref_dq[sl]) < 2e-2  # bf16 query gather
        assert np.linalg.norm(o["dk"] - 0.5 * ref_dk[sl]) / np.linalg.norm(0.5 * ref_dk[sl]) < 2e-2
    assert out[0]["loss"] == out[1]["loss"]


# ======================================================================================================= ReLIC (KL over the global batch axis)
class EmulatedRelicKlStages:
    """torch-fp64 stand-in for ssvb_relic_kl_dist_dots / _dist_reduce / ssvb_relic_kl_bwd (test-only)."""

    def alloc_saved(self, n, dev):
        return {}

    def dots(self, zi, zj, zo, normalize, temperature, saved, ab_local):
        def nrm(z):
            z = z.double()
            den = z.norm(dim=1, keepdim=True).clamp_min(1e-12) if normalize else torch.ones(z.shape[0], 1, dtype=torch.float64)
            return z / den, den
        (ih, idn), (jh, jdn), (oh, odn) = nrm(zi), nrm(zj), nrm(zo)
        a, b = (ih * oh).sum(1) / temperature, (jh * oh).sum(1) / temperature
        saved.update(ih=ih, jh=jh, oh=oh, idn=idn, jdn=jdn, odn=odn, a=a, b=b)
        ab_local[0].copy_(a.float())
        ab_local[1].copy_(b.float())

    def reduce(self, ab_all, world, n_local, alpha, saved, kl):
        a = ab_all[:, 0, :].reshape(-1).double()      # rank-order concatenation = the global batch axis
        b = ab_all[:, 1, :].reshape(-1).double()
        lse_a, lse_b = torch.logsumexp(a, 0), torch.logsumexp(b, 0)
        p, lq = torch.exp(a - lse_a), b - lse_b
        q = torch.exp(lq)
        klv = (q * (lq - p)).sum()
        saved.update(lse_a=lse_a, lse_b=lse_b, spq=(p * q).sum(), kl=klv)
        kl.copy_((alpha * klv).float())

    def bwd(self, zi, zj, zo, normalize, temperature, alpha, grad_out, saved, dzi, dzj, dzo):
        s = saved
        p, lq = torch.exp(s["a"] - s["lse_a"]), s["b"] - s["lse_b"]
        q = torch.exp(lq)
        g = grad_out.double() * alpha / temperature
        da = -(p * q - p * s["spq"]) * g
        db = (q * (lq - p) + q - q * (s["kl"] + 1.0)) * g
        dih, djh = da[:, None] * s["oh"], db[:, None] * s["oh"]
        doh = da[:, None] * s["ih"] + db[:, None] * s["jh"]

        def back(dh, h, den):
            return (dh - (dh * h).sum(1, keepdim=True) * h) / den if normalize else dh
        dzi.add_(back(dih, s["ih"], s["idn"]).float())
        dzj.add_(back(djh, s["jh"], s["jdn"]).float())
        dzo.copy_(back(doh, s["oh"], s["odn"]).float())


def _relic_worker(rank, world, port, n_local, d, normalize, tau, alpha, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "self-supervised-vision_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ssv_b200.dist import DistributedRelicLoss
    g = torch.Generator().manual_seed(300 + rank)
    zi, zj, zo = (torch.randn(n_local, d, generator=g, requires_grad=True) for _ in range(3))
    fn = DistributedRelicLoss(normalize, tau, alpha, stages=EmulatedStages(), kl_stages=EmulatedRelicKlStages())
    loss = fn(zi, zj, zo)
    loss.backward()
    out[rank] = (loss.item(),) + tuple(t.grad.numpy().copy() for t in (zi, zj, zo)) + \
        tuple(t.detach().numpy().copy() for t in (zi, zj, zo))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("normalize,tau,alpha", [(True, 1.0, 0.5), (True, 0.5, 2.0)])
def test_distributed_relic_matches_global_oracle(normalize, tau, alpha):
    """DistributedRelicLoss under gloo (world 2): the KL's batch-axis softmaxes must span BOTH ranks' rows
    (reference utils/losses.py:196-200 on the concatenation), each rank gets its own gradient rows."""
    from oracle import ssl_oracle as O
    world, n_local, d = 2, 20, 16
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000) + 7
    mp.spawn(_relic_worker, args=(world, port, n_local, d, normalize, tau, alpha, out), nprocs=world, join=True)
    zi, zj, zo = (np.concatenate([out[r][4 + k] for r in range(world)]) for k in range(3))
    ref = O.relic(zi, zj, zo, normalize, tau, alpha)
    for r in range(world):
        sl = slice(r * n_local, (r + 1) * n_local)
        assert abs(out[r][0] - ref[0]) / abs(ref[0]) < 1e-4
        for k in range(3):
            g, rg = out[r][1 + k], ref[1 + k][sl]
            assert np.linalg.norm(g - rg) / np.linalg.norm(rg) < 2e-2, f"rank {r} grad {k}"
    assert out[0][0] == out[1][0]
