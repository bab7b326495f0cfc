"""Global-batch NT-Xent over one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The reference has no multi-GPU path; semantics: SimclrLoss evaluated on the concatenation
of all ranks' (zi, zj); every rank gets the gradient rows of its own inputs, no 1/world rescale.

Row sharding: rank r owns the 2L rows of its local batch (L = per-rank batch) and computes their
similarity rows against ALL columns.  One exchange each way:
  forward : all-gather of the bf16 normalised rows (2L x dpad per rank), all-reduce of the scalar loss,
            all-gather of the per-row LSE (2L floats per rank) - issued in forward so it is done before backward;
  backward: nothing.  W_ab = P_ab + P_ba only needs s_ab, lse_a, lse_b, so each rank produces the complete
            gradient of its own rows and no gradient reduce-scatter exists.
The gathered matrices are rank-major (see include/ssv_b200.h), so each collective is a single call on a
contiguous slot.  The three compute stages go through `stages` (default: the CUDA library); tests inject a
CPU emulation of the stages to exercise this orchestration under gloo.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _cabi as C


class CudaStages:
    """The product path: hand-written kernels behind the C ABI."""

    def dpad(self, d):
        return int(C.lib().ssvb_ntxent_dpad(d))

    def mpad(self, n_global):
        return int(C.lib().ssvb_ntxent_mpad(n_global))

    def prep(self, zi, zj, normalize, world, rank, zhat_all, inv_local, pos_local):
        n, d = zi.shape
        C.check(C.lib().ssvb_ntxent_dist_prep(C.ptr(zi), C.ptr(zj), n, d, zi.stride(0), zj.stride(0), normalize, world,
                                              rank, C.ptr(zhat_all), C.ptr(inv_local), C.ptr(pos_local),
                                              C.stream_ptr(zi.device)), "ssvb_ntxent_dist_prep")

    def rows_fwd(self, zhat_all, world, rank, n_local, d, normalize, temperature, pos_local, stat_local, loss_sum):
        L = C.lib()
        dev = zhat_all.device
        ws_bytes = L.ssvb_ntxent_dist_workspace_bytes(world, n_local, d)
        ws = C.byte_buffer(ws_bytes, dev)
        C.check(L.ssvb_ntxent_dist_rows_fwd(C.ptr(zhat_all), world, rank, n_local, d, normalize, temperature,
                                            C.ptr(pos_local), C.ptr(stat_local), C.ptr(loss_sum), C.ptr(ws), ws_bytes,
                                            C.stream_ptr(dev)), "ssvb_ntxent_dist_rows_fwd")

    def rows_bwd(self, zi, zj, normalize, temperature, world, rank, zhat_all, stat_all, inv_local, grad_out, dzi, dzj):
        L = C.lib()
        n, d = zi.shape
        dev = zi.device
        ws_bytes = L.ssvb_ntxent_dist_workspace_bytes(world, n, d)
        ws = C.byte_buffer(ws_bytes, dev)
        C.check(L.ssvb_ntxent_dist_rows_bwd(C.ptr(zi), C.ptr(zj), n, d, zi.stride(0), zj.stride(0), normalize,
                                            temperature, world, rank, C.ptr(zhat_all), C.ptr(stat_all), C.ptr(inv_local),
                                            C.ptr(grad_out), C.ptr(dzi), C.ptr(dzj), dzi.stride(0), dzj.stride(0),
                                            C.ptr(ws), ws_bytes, C.stream_ptr(dev)), "ssvb_ntxent_dist_rows_bwd")


def _gather_slots(full, slot, group, inplace):
    """all-gather `slot` (this rank's contiguous rows of `full`) into `full` (world * slot rows)."""
    if inplace:  # NCCL: sendbuf == recvbuf + rank * count is the in-place form
        dist.all_gather_into_tensor(full, slot, group=group)
    else:        # gloo (tests): no aliasing between input and output
        parts = [torch.empty_like(slot) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, slot.clone(), group=group)
        full.copy_(torch.cat(parts, 0))


class _NtxentDistFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, zi, zj, normalize, temperature, group, stages):
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        cuda = isinstance(stages, CudaStages)
        if cuda:
            C.require_cuda(zi, zj)
            xi, xj = C.as_f32_rows(zi), C.as_f32_rows(zj)
        else:
            xi, xj = zi.detach().float().contiguous(), zj.detach().float().contiguous()
        n, d = xi.shape
        dev = xi.device
        m = 2 * n * world
        mpad, dpad = stages.mpad(n * world), stages.dpad(d)
        norm = int(bool(normalize))
        zhat_all = torch.empty(mpad, dpad, dtype=torch.bfloat16, device=dev)
        inv_local = torch.empty(2 * n, dtype=torch.float32, device=dev)
        pos_local = torch.empty(2 * n, dtype=torch.float32, device=dev)
        stat_all = torch.empty(m, dtype=torch.float32, device=dev)
        loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
        stages.prep(xi, xj, norm, world, rank, zhat_all, inv_local, pos_local)
        my = slice(rank * 2 * n, (rank + 1) * 2 * n)
        if world > 1:
            _gather_slots(zhat_all[:m], zhat_all[my], group, inplace=cuda)
        stages.rows_fwd(zhat_all, world, rank, n, d, norm, float(temperature), pos_local, stat_all[my], loss_sum)
        if world > 1:
            _gather_slots(stat_all, stat_all[my], group, inplace=cuda)
            dist.all_reduce(loss_sum, group=group)
        ctx.save_for_backward(xi, xj, zhat_all, stat_all, inv_local)
        ctx.cfg = (norm, float(temperature), world, rank, stages, zi.dtype, zj.dtype)
        return loss_sum / m

    @staticmethod
    def backward(ctx, grad_out):
        xi, xj, zhat_all, stat_all, inv_local = ctx.saved_tensors
        norm, temperature, world, rank, stages, dti, dtj = ctx.cfg
        go = grad_out.to(torch.float32).contiguous()
        dzi, dzj = torch.empty_like(xi), torch.empty_like(xj)
        stages.rows_bwd(xi, xj, norm, temperature, world, rank, zhat_all, stat_all, inv_local, go, dzi, dzj)
        return dzi.to(dti), dzj.to(dtj), None, None, None, None


class DistributedSimclrLoss(nn.Module):
    """SimclrLoss over the global batch of a process group: same ctor kwargs as the reference's
    SimclrLoss (utils/losses.py:10-13) plus an optional process group."""

    def __init__(self, normalize=False, temperature=1.0, group=None, stages=None):
        super().__init__()
        self.normalize = normalize
        self.temperature = temperature
        self.group = group
        self.stages = stages if stages is not None else CudaStages()

    def forward(self, zi, zj):
        return _NtxentDistFn.apply(zi, zj, self.normalize, self.temperature, self.group, self.stages)
