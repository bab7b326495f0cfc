"""Global-batch NT-Xent over one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The reference has no multi-GPU path; semantics: SimclrLoss evaluated on the concatenation
of all ranks' (zi, zj); every rank gets the gradient rows of its own inputs, no 1/world rescale.

Row sharding: rank r owns the 2L rows of its local batch (L = per-rank batch) and computes their
similarity rows against ALL columns.  One exchange each way:
  transport "p2p" (default when torch symmetric memory works): the normalise kernel and the LSE-finalize kernel
            store their rank's slot directly into EVERY peer's gather buffer over NVLink (fused compute +
            all-gather), a symmetric-memory barrier publishes it - no NCCL on the data path;
  transport "nccl":
  forward : all-gather of the bf16 normalised rows (2L x dpad per rank), then ONE all-gather of
            [per-row LSE | per-row loss term] (4L floats per rank) that serves both the backward's column
            statistics and the global loss (summed locally, identical on every rank: no all-reduce);
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
        return C.cached_size("ssvb_ntxent_dpad", d)

    def mpad(self, n_global):
        return C.cached_size("ssvb_ntxent_mpad", n_global)

    def prep(self, zi, zj, normalize, world, rank, zhat_all, inv_local, pos_local):
        n, d = zi.shape
        C.check(C.lib().ssvb_ntxent_dist_prep(C.ptr(zi), C.ptr(zj), n, d, zi.stride(0), zj.stride(0), normalize, world,
                                              rank, C.ptr(zhat_all), C.ptr(inv_local), C.ptr(pos_local),
                                              C.stream_ptr(zi.device)), "ssvb_ntxent_dist_prep")

    def rows_fwd(self, zhat_all, world, rank, n_local, d, normalize, temperature, pos_local, stat_local, loss_sum):
        L = C.lib()
        dev = zhat_all.device
        ws_bytes = C.cached_size("ssvb_ntxent_dist_workspace_bytes", world, n_local, d)
        ws = C.workspace("ntxent_dist", ws_bytes, dev)
        C.check(L.ssvb_ntxent_dist_rows_fwd(C.ptr(zhat_all), world, rank, n_local, d, normalize, temperature,
                                            C.ptr(pos_local), C.ptr(stat_local), C.ptr(loss_sum), C.ptr(ws), ws_bytes,
                                            C.stream_ptr(dev)), "ssvb_ntxent_dist_rows_fwd")

    def rows_bwd(self, zi, zj, normalize, temperature, world, rank, zhat_all, stat_all, inv_local, grad_out, dzi, dzj):
        L = C.lib()
        n, d = zi.shape
        dev = zi.device
        ws_bytes = C.cached_size("ssvb_ntxent_dist_workspace_bytes", world, n, d)
        ws = C.workspace("ntxent_dist", ws_bytes, dev)
        C.check(L.ssvb_ntxent_dist_rows_bwd(C.ptr(zi), C.ptr(zj), n, d, zi.stride(0), zj.stride(0), normalize,
                                            temperature, world, rank, C.ptr(zhat_all), C.ptr(stat_all), C.ptr(inv_local),
                                            C.ptr(grad_out), C.ptr(dzi), C.ptr(dzj), dzi.stride(0), dzj.stride(0),
                                            C.ptr(ws), ws_bytes, C.stream_ptr(dev)), "ssvb_ntxent_dist_rows_bwd")


    # ---- fused compute + all-gather over NVLink peer memory (no NCCL on the data path)
    def prep_push(self, zi, zj, normalize, world, rank, peer_zhat_dev, inv_local, pos_local):
        n, d = zi.shape
        C.check(C.lib().ssvb_ntxent_dist_prep_push(C.ptr(zi), C.ptr(zj), n, d, zi.stride(0), zj.stride(0), normalize,
                                                   world, rank, C.c_void_p(peer_zhat_dev), C.ptr(inv_local),
                                                   C.ptr(pos_local), C.stream_ptr(zi.device)), "ssvb_ntxent_dist_prep_push")

    def rows_fwd_push(self, zhat_all, world, rank, n_local, d, normalize, temperature, pos_local, peer_stat_dev, loss_sum):
        L = C.lib()
        dev = zhat_all.device
        ws_bytes = C.cached_size("ssvb_ntxent_dist_workspace_bytes", world, n_local, d)
        ws = C.workspace("ntxent_dist", ws_bytes, dev)
        C.check(L.ssvb_ntxent_dist_rows_fwd_push(C.ptr(zhat_all), world, rank, n_local, d, normalize, temperature,
                                                 C.ptr(pos_local), C.c_void_p(peer_stat_dev), C.ptr(loss_sum), C.ptr(ws),
                                                 ws_bytes, C.stream_ptr(dev)), "ssvb_ntxent_dist_rows_fwd_push")

    def dist_loss(self, stat_all, world, n_local, loss):
        C.check(C.lib().ssvb_ntxent_dist_loss(C.ptr(stat_all), world, n_local, C.ptr(loss),
                                              C.stream_ptr(stat_all.device)), "ssvb_ntxent_dist_loss")


class _PeerTransport:
    """Double-buffered symmetric (peer-mapped) gather buffers for one (group, shape): every rank's kernels store
    their slot straight into all peers' buffers over NVLink; a symmetric-memory barrier publishes the data.
    Double buffering makes a pre-push barrier unnecessary (a rank can never be two pushes ahead of a peer: it must
    pass the peer's previous post-push barrier first), and the gathered data is copied into a private tensor before
    use, so the autograd graph never references a buffer a later forward may overwrite."""

    _cache = {}

    def __init__(self, group, world, mpad, dpad, n_local, dev):
        import torch.distributed._symmetric_memory as symm_mem
        self.bufs = []
        for _ in range(2):
            z = symm_mem.empty(mpad * dpad, dtype=torch.bfloat16, device=dev)
            z.zero_()  # padding rows stay zero for ever (nobody writes them)
            hz = symm_mem.rendezvous(z, group)
            st = symm_mem.empty(world * 4 * n_local, dtype=torch.float32, device=dev)
            hs = symm_mem.rendezvous(st, group)
            hz.barrier()
            self.bufs.append((z, hz, st, hs))
        self.parity = 0

    @classmethod
    def get(cls, group, world, mpad, dpad, n_local, dev):
        g = group if group is not None else dist.group.WORLD
        key = (id(g), world, mpad, dpad, n_local, dev.index)
        if key not in cls._cache:
            cls._cache[key] = cls(g, world, mpad, dpad, n_local, dev)
        return cls._cache[key]

    def next(self):
        self.parity ^= 1
        return self.bufs[self.parity]


_P2P_BROKEN = False


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
    def forward(ctx, zi, zj, normalize, temperature, group, stages, transport):
        global _P2P_BROKEN
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
        mpad, dpad = stages.mpad(n * world), stages.dpad(d)  # (memoised for the CUDA stages)
        norm = int(bool(normalize))
        inv_local = torch.empty(2 * n, dtype=torch.float32, device=dev)
        pos_local = torch.empty(2 * n, dtype=torch.float32, device=dev)
        peer = None
        if cuda and world > 1 and transport in ("auto", "p2p") and not _P2P_BROKEN:
            try:
                peer = _PeerTransport.get(group, world, mpad, dpad, n, dev)
            except Exception:  # symmetric memory unavailable on this system: NCCL collectives instead
                if transport == "p2p":
                    raise
                _P2P_BROKEN = True
        if peer is not None:
            zbuf, hz, sbuf, hs = peer.next()
            loss = torch.empty((), dtype=torch.float32, device=dev)
            stages.prep_push(xi, xj, norm, world, rank, hz.buffer_ptrs_dev, inv_local, pos_local)
            hz.barrier()
            zhat_all = zbuf.view(mpad, dpad).clone()
            stages.rows_fwd_push(zhat_all, world, rank, n, d, norm, float(temperature), pos_local, hs.buffer_ptrs_dev, loss)
            hs.barrier()
            stat_all = sbuf.view(world, 2, 2 * n).clone()
            stages.dist_loss(stat_all, world, n, loss)
            ctx.save_for_backward(xi, xj, zhat_all, stat_all, inv_local)
            ctx.cfg = (norm, float(temperature), world, rank, stages, zi.dtype, zj.dtype)
            return loss
        zhat_all = torch.empty(mpad, dpad, dtype=torch.bfloat16, device=dev)
        # per rank: [lse2 (2L) | per-row loss term (2L)] -> one all-gather serves the backward AND the global loss
        stat_all = torch.empty(world, 2, 2 * n, dtype=torch.float32, device=dev)
        loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
        stages.prep(xi, xj, norm, world, rank, zhat_all, inv_local, pos_local)
        my = slice(rank * 2 * n, (rank + 1) * 2 * n)
        if world > 1:
            _gather_slots(zhat_all[:m], zhat_all[my], group, inplace=cuda)
        stages.rows_fwd(zhat_all, world, rank, n, d, norm, float(temperature), pos_local, stat_all[rank], loss_sum)
        if world > 1:
            _gather_slots(stat_all.view(world * 2, 2 * n), stat_all[rank], group, inplace=cuda)
            if cuda:
                loss = torch.empty((), dtype=torch.float32, device=dev)
                stages.dist_loss(stat_all, world, n, loss)
            else:
                loss = stat_all[:, 1, :].sum() / m   # identical on every rank (same data, same reduction)
        else:
            loss = loss_sum / m
        ctx.save_for_backward(xi, xj, zhat_all, stat_all, inv_local)
        ctx.cfg = (norm, float(temperature), world, rank, stages, zi.dtype, zj.dtype)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        xi, xj, zhat_all, stat_all, inv_local = ctx.saved_tensors
        norm, temperature, world, rank, stages, dti, dtj = ctx.cfg
        go = C.f32_scalar(grad_out)
        dzi, dzj = torch.empty_like(xi), torch.empty_like(xj)
        stages.rows_bwd(xi, xj, norm, temperature, world, rank, zhat_all, stat_all, inv_local, go, dzi, dzj)
        return dzi.to(dti), dzj.to(dtj), None, None, None, None, None


class DistributedSimclrLoss(nn.Module):
    """SimclrLoss over the global batch of a process group: same ctor kwargs as the reference's
    SimclrLoss (utils/losses.py:10-13) plus an optional process group."""

    def __init__(self, normalize=False, temperature=1.0, group=None, stages=None, transport="auto"):
        """transport: "p2p" = kernels store into all peers' buffers over NVLink (torch symmetric memory) + barriers;
        "nccl" = two all-gathers; "auto" = p2p when symmetric memory is available, else nccl."""
        super().__init__()
        self.normalize = normalize
        self.temperature = temperature
        self.group = group
        self.stages = stages if stages is not None else CudaStages()
        self.transport = transport

    def forward(self, zi, zj):
        return _NtxentDistFn.apply(zi, zj, self.normalize, self.temperature, self.group, self.stages, self.transport)
