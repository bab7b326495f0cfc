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

    def prep(self, zi, zj, normalize, world, rank, zhat_all, inv_local, pos_local):
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
