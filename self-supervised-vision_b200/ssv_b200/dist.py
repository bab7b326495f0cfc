"""Global-batch NT-Xent over one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

The reference has no multi-GPU path; semantics: SimclrLoss evaluated on the concatenation
of all ranks' (zi, zj); every rank gets the gradient rows of its own inputs, no 1/world rescale.

Row sharding: rank r owns the 2L rows of its local batch (L = per-rank batch) and computes their
similarity rows against ALL columns.  One exchange each way:
  transport "p2p" (default when torch symmetric memory works): fused compute + all-gather over NVLink peer memory.
            Every rank owns one symmetric arena; the normalise kernel stores its rows into EVERY arena (unicast peer
            stores, or ONE multicast store through the NVSwitch when the allocation has a multicast mapping) and the
            LSE-finalize kernel does the same with its [lse | term] block; completion is published with per-rank
            generation flags that the consuming kernels poll on the device - no NCCL call and no host-issued barrier
            on the data path, and the buffers are double-buffered by generation parity so no pre-push barrier exists;
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
from torch.autograd.function import once_differentiable

from . import _cabi as C


class CudaStages:
    """The product path: hand-written kernels behind the C ABI."""

    def dpad(self, d):
        return C.cached_size("ssvb_ntxent_dpad", d)

    def mpad(self, n_global):
        return C.cached_size("ssvb_ntxent_mpad", n_global)

    def prep(self, zi, zj, normalize, temperature, world, rank, zhat_all, inv_local, pos_local):
        n, d = zi.shape
        C.check(C.lib().ssvb_ntxent_dist_prep(C.ptr(zi), C.ptr(zj), n, d, zi.stride(0), zj.stride(0), normalize,
                                              temperature, world,
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


    # ---- fused compute + all-gather over NVLink peer memory (generation flags, no NCCL / barrier on the data path)
    def p2p_prep_push(self, zi, zj, normalize, temperature, world, rank, arena, gen, inv_local, pos_local):
        n, d = zi.shape
        C.check(C.lib().ssvb_ntxent_p2p_prep_push(C.ptr(zi), C.ptr(zj), n, d, zi.stride(0), zj.stride(0), normalize,
                                                  temperature, world, rank, C.c_void_p(arena.local_ptr),
                                                  C.c_void_p(arena.peers_dev), C.c_void_p(arena.multicast_ptr or None),
                                                  gen, C.ptr(inv_local), C.ptr(pos_local), C.stream_ptr(zi.device)),
                "ssvb_ntxent_p2p_prep_push")

    def p2p_wait_copy(self, arena, world, rank, n_local, d, gen, zhat_all):
        C.check(C.lib().ssvb_ntxent_p2p_wait_copy(C.c_void_p(arena.local_ptr), world, rank, n_local, d, gen,
                                                  C.ptr(zhat_all), C.stream_ptr(zhat_all.device)),
                "ssvb_ntxent_p2p_wait_copy")

    def _dist_ws(self, world, n_local, d, dev):
        ws_bytes = C.cached_size("ssvb_ntxent_dist_workspace_bytes", world, n_local, d)
        return C.workspace("ntxent_dist", ws_bytes, dev), ws_bytes

    def p2p_rows_fwd(self, zhat_all, world, rank, n_local, d, normalize, temperature, pos_local, arena, gen, loss_sum):
        dev = zhat_all.device
        ws, ws_bytes = self._dist_ws(world, n_local, d, dev)
        C.check(C.lib().ssvb_ntxent_p2p_rows_fwd(C.ptr(zhat_all), world, rank, n_local, d, normalize, temperature,
                                                 C.ptr(pos_local), C.c_void_p(arena.local_ptr), C.c_void_p(arena.peers_dev),
                                                 gen, C.ptr(loss_sum),
                                                 C.ptr(ws), ws_bytes, C.stream_ptr(dev)), "ssvb_ntxent_p2p_rows_fwd")

    def p2p_stat_loss(self, arena, world, rank, n_local, d, normalize, temperature, gen, colstat, loss):
        dev = colstat.device
        ws, ws_bytes = self._dist_ws(world, n_local, d, dev)
        C.check(C.lib().ssvb_ntxent_p2p_stat_loss(C.c_void_p(arena.local_ptr), world, rank, n_local, d, normalize,
                                                  temperature, gen, C.ptr(colstat), C.ptr(loss), C.ptr(ws), ws_bytes,
                                                  C.stream_ptr(dev)), "ssvb_ntxent_p2p_stat_loss")

    def p2p_forward(self, zi, zj, normalize, temperature, world, rank, arena, gen, zhat_all, aux_ptr, n_aux, loss):
        """All four forward stages in one C call.  `aux_ptr` = base address of one fp32 scratch tensor laid out as
        [inv_norm (2L) | pos (2L) | colstat (mpad) | loss_sum | pad] (`n_aux` = (2L, mpad))."""
        n, d = zi.shape
        dev = zi.device
        ws, ws_bytes = self._dist_ws(world, n, d, dev)
        two_l, mpad = n_aux
        p_inv, p_pos = aux_ptr, aux_ptr + 4 * two_l
        p_col = aux_ptr + 8 * two_l
        p_sum = p_col + 4 * mpad
        C.check(C.lib().ssvb_ntxent_p2p_forward(C.ptr(zi), C.ptr(zj), n, d, zi.stride(0), zj.stride(0), normalize,
                                                temperature, world, rank, C.c_void_p(arena.local_ptr),
                                                C.c_void_p(arena.peers_dev), C.c_void_p(arena.multicast_ptr or None), gen,
                                                C.ptr(zhat_all), C.c_void_p(p_inv), C.c_void_p(p_pos), C.c_void_p(p_col),
                                                C.c_void_p(p_sum), C.ptr(loss), C.ptr(ws), ws_bytes,
                                                C.stream_ptr(dev)), "ssvb_ntxent_p2p_forward")

    def p2p_rows_bwd(self, zi, zj, normalize, temperature, world, rank, zhat_all, colstat, inv_local, grad_out, dzi, dzj):
        n, d = zi.shape
        dev = zi.device
        ws, ws_bytes = self._dist_ws(world, n, d, dev)
        as_ptr = lambda x: C.c_void_p(x) if isinstance(x, int) else C.ptr(x)  # noqa: E731  (tensor or raw device address)
        C.check(C.lib().ssvb_ntxent_p2p_rows_bwd(C.ptr(zi), C.ptr(zj), n, d, zi.stride(0), zj.stride(0), normalize,
                                                 temperature, world, rank, C.ptr(zhat_all), as_ptr(colstat),
                                                 as_ptr(inv_local), C.ptr(grad_out), C.ptr(dzi), C.ptr(dzj),
                                                 dzi.stride(0), dzj.stride(0), C.ptr(ws), ws_bytes, C.stream_ptr(dev)),
                "ssvb_ntxent_p2p_rows_bwd")

    def dist_loss(self, stat_all, world, n_local, loss):
        C.check(C.lib().ssvb_ntxent_dist_loss(C.ptr(stat_all), world, n_local, C.ptr(loss),
                                              C.stream_ptr(stat_all.device)), "ssvb_ntxent_dist_loss")


class _PeerArena:
    """One symmetric (peer-mapped) arena per (group, shape) holding the double-buffered gather buffers and the
    generation flags of the NVLink transport (layout: csrc/ntxent.cu `ArenaLayout`).  Created collectively: zeroed,
    then ONE group barrier so that no peer can push before the flags are cleared; after that the protocol needs no
    barrier (see include/ssv_b200.h).  `gen` counts the forwards of this arena and is identical on every rank."""

    _cache = {}

    def __init__(self, group, world, n_local, d, dev, multicast):
        import torch.distributed._symmetric_memory as symm_mem
        nbytes = C.cached_size("ssvb_ntxent_p2p_arena_bytes", world, n_local, d)
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.hdl.barrier()
        self.local_ptr = self.buf.data_ptr()
        self.peers_dev = self.hdl.buffer_ptrs_dev
        mc = getattr(self.hdl, "multicast_ptr", 0) or 0
        self.multicast_ptr = mc if multicast else 0
        self.gen = 0

    @classmethod
    def get(cls, group, world, n_local, d, dev, multicast=True):
        g = group if group is not None else dist.group.WORLD
        key = (id(g), world, n_local, d, dev.index, bool(multicast))
        if key not in cls._cache:
            cls._cache[key] = cls(g, world, n_local, d, dev, multicast)
        return cls._cache[key]

    def next_gen(self):
        self.gen += 1
        return self.gen


_P2P_STATE = {}   # id(group) -> True (peer arenas work on EVERY rank) / False (some rank failed: NCCL everywhere)


def _p2p_available(group, world, n_local, d, dev, multicast, required):
    """The transport must be the same on every rank (mixing peer flags with NCCL collectives would deadlock), so the
    outcome of the arena set-up is agreed on collectively: MIN over the ranks of a success flag."""
    g = group if group is not None else dist.group.WORLD
    key = (id(g), world, n_local, d, dev.index, bool(multicast))
    if key in _P2P_STATE:
        return _P2P_STATE[key]
    ok, err = 1, None
    try:
        _PeerArena.get(group, world, n_local, d, dev, multicast)
    except Exception as e:  # noqa: BLE001  (symmetric memory unavailable on this system)
        ok, err = 0, e
    flag = torch.tensor([ok], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    good = bool(flag.item())
    _P2P_STATE[key] = good
    if not good and required:
        raise RuntimeError(f"transport='p2p' requested but peer-memory arenas are unavailable on some rank: {err!r}")
    return good


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
        use_p2p = False
        multicast = transport != "p2p-unicast"
        if cuda and world > 1 and transport in ("auto", "p2p", "p2p-unicast") and (d <= 128 or transport != "auto"):
            # (the peer-push kernels cover d <= 128; "auto" takes the NCCL transport for 128 < d <= 256)
            use_p2p = _p2p_available(group, world, n, d, dev, multicast, required=transport != "auto")
        if use_p2p:
            arena = _PeerArena.get(group, world, n, d, dev, multicast)
            gen = arena.next_gen()
            tau = float(temperature)
            zhat_all = torch.empty(mpad, dpad, dtype=torch.bfloat16, device=dev)
            # one fp32 scratch tensor for everything small: [inv_norm 2L | pos 2L | colstat mpad | loss_sum | loss]
            # (each torch.empty costs ~2 us of host time and the whole 8-GPU step is ~0.4 ms)
            two_l = 2 * n
            aux = torch.empty(2 * two_l + mpad + 2, dtype=torch.float32, device=dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)   # (an output must not be a view of a saved tensor)
            stages.p2p_forward(xi, xj, norm, tau, world, rank, arena, gen, zhat_all, aux.data_ptr(), (two_l, mpad), loss)
            # everything the backward needs is private memory: the arena is never read again after this forward
            ctx.save_for_backward(xi, xj, zhat_all, aux)
            ctx.cfg = (norm, tau, world, rank, stages, zi.dtype, zj.dtype)
            ctx.p2p = (two_l, mpad)
            return loss
        inv_local = torch.empty(2 * n, dtype=torch.float32, device=dev)
        pos_local = torch.empty(2 * n, dtype=torch.float32, device=dev)
        zhat_all = torch.empty(mpad, dpad, dtype=torch.bfloat16, device=dev)
        # per rank: [lse2 (2L) | per-row loss term (2L)] -> one all-gather serves the backward AND the global loss
        stat_all = torch.empty(world, 2, 2 * n, dtype=torch.float32, device=dev)
        loss_sum = torch.zeros((), dtype=torch.float32, device=dev)
        stages.prep(xi, xj, norm, float(temperature), world, rank, zhat_all, inv_local, pos_local)
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
        ctx.p2p = None
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        norm, temperature, world, rank, stages, dti, dtj = ctx.cfg
        go = C.f32_scalar(grad_out)
        if ctx.p2p is not None:
            xi, xj, zhat_all, aux = ctx.saved_tensors
            two_l, mpad = ctx.p2p
            dzi, dzj = torch.empty_like(xi), torch.empty_like(xj)
            base = aux.data_ptr()   # raw addresses instead of two view objects: [inv_norm | pos | colstat | ...]
            stages.p2p_rows_bwd(xi, xj, norm, temperature, world, rank, zhat_all, base + 8 * two_l, base, go, dzi, dzj)
        else:
            xi, xj, zhat_all, stat_all, inv_local = ctx.saved_tensors
            dzi, dzj = torch.empty_like(xi), torch.empty_like(xj)
            stages.rows_bwd(xi, xj, norm, temperature, world, rank, zhat_all, stat_all, inv_local, go, dzi, dzj)
        return dzi.to(dti), dzj.to(dtj), None, None, None, None, None


class DistributedSimclrLoss(nn.Module):
    """SimclrLoss over the global batch of a process group: same ctor kwargs as the reference's
    SimclrLoss (utils/losses.py:10-13) plus an optional process group."""

    def __init__(self, normalize=False, temperature=1.0, group=None, stages=None, transport="auto",
                 retain_gathered=True):
        """transport: "p2p" = kernels store into all peers' arenas over NVLink (torch symmetric memory; one multicast
        store through the NVSwitch when available) and poll generation flags on the device; "p2p-unicast" = the same
        with per-peer stores only; "nccl" = two all-gathers; "auto" = p2p when the arenas can be set up on EVERY rank
        (agreed collectively), else nccl.
        retain_gathered: kept for API compatibility; the backward only ever reads private copies, so any
        forward / backward interleaving is valid."""
        super().__init__()
        self.retain_gathered = retain_gathered
        self.normalize = normalize
        self.temperature = temperature
        self.group = group
        self.stages = stages if stages is not None else CudaStages()
        self.transport = transport

    def forward(self, zi, zj):
        return _NtxentDistFn.apply(zi, zj, self.normalize, self.temperature, self.group, self.stages, self.transport)


# ======================================================================================================= ReLIC
class RelicKlCudaStages:
    """The KL term of the distributed ReLIC loss: `ssvb_relic_kl_dist_dots / _dist_reduce` + `ssvb_relic_kl_bwd`."""

    def alloc_saved(self, n, dev):
        return C.byte_buffer(C.cached_size("ssvb_relic_kl_saved_bytes", n), dev)

    def dots(self, zi, zj, zo, normalize, temperature, saved, ab_local):
        n, d = zi.shape
        C.check(C.lib().ssvb_relic_kl_dist_dots(C.ptr(zi), C.ptr(zj), C.ptr(zo), n, d, zi.stride(0), zj.stride(0),
                                                zo.stride(0), normalize, temperature, C.ptr(saved), C.ptr(ab_local),
                                                C.stream_ptr(zi.device)), "ssvb_relic_kl_dist_dots")

    def reduce(self, ab_all, world, n_local, alpha, saved, kl):
        C.check(C.lib().ssvb_relic_kl_dist_reduce(C.ptr(ab_all), world, n_local, alpha, C.ptr(saved), C.ptr(kl),
                                                  C.stream_ptr(ab_all.device)), "ssvb_relic_kl_dist_reduce")

    def bwd(self, zi, zj, zo, normalize, temperature, alpha, grad_out, saved, dzi, dzj, dzo):
        n, d = zi.shape
        C.check(C.lib().ssvb_relic_kl_bwd(C.ptr(zi), C.ptr(zj), C.ptr(zo), n, d, zi.stride(0), zj.stride(0),
                                          zo.stride(0), normalize, temperature, alpha, C.ptr(grad_out), C.ptr(saved),
                                          C.ptr(dzi), C.ptr(dzj), C.ptr(dzo), dzi.stride(0), dzj.stride(0),
                                          dzo.stride(0), C.stream_ptr(zi.device)), "ssvb_relic_kl_bwd")


class _RelicKlDistFn(torch.autograd.Function):
    """alpha * KL_quirk of RelicLoss (utils/losses.py:196-200) with the two softmaxes taken over the GLOBAL batch axis:
    one all-gather of this rank's two N_local-vectors of logits; no exchange in backward (the global softmax
    statistics are identical on every rank and each rank differentiates its own rows)."""

    @staticmethod
    def forward(ctx, zi, zj, zo, normalize, temperature, alpha, group, stages):
        world, rank = _world_rank(group)
        cuda = isinstance(stages, RelicKlCudaStages)
        if cuda:
            C.require_cuda(zi, zj, zo)
            xi, xj, xo = C.as_f32_rows(zi), C.as_f32_rows(zj), C.as_f32_rows(zo)
        else:
            xi, xj, xo = (t.detach().float().contiguous() for t in (zi, zj, zo))
        n = xi.shape[0]
        dev = xi.device
        norm = int(bool(normalize))
        saved = stages.alloc_saved(n, dev)
        ab_all = torch.empty(world, 2, n, dtype=torch.float32, device=dev)
        stages.dots(xi, xj, xo, norm, float(temperature), saved, ab_all[rank])
        if world > 1:
            _gather_slots(ab_all.view(world * 2, n), ab_all[rank], group, inplace=cuda and _is_nccl(group))
        kl = torch.empty((), dtype=torch.float32, device=dev)
        stages.reduce(ab_all, world, n, float(alpha), saved, kl)
        if torch.is_tensor(saved):
            ctx.save_for_backward(xi, xj, xo, saved)
        else:   # emulated stages (tests) keep a host-side object
            ctx.save_for_backward(xi, xj, xo)
            ctx.saved_obj = saved
        ctx.cfg = (norm, float(temperature), float(alpha), stages, zi.dtype, zj.dtype, zo.dtype)
        return kl

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        if len(ctx.saved_tensors) == 4:
            xi, xj, xo, saved = ctx.saved_tensors
        else:
            (xi, xj, xo), saved = ctx.saved_tensors, ctx.saved_obj
        norm, temperature, alpha, stages, dti, dtj, dto = ctx.cfg
        go = C.f32_scalar(grad_out)
        # ssvb_relic_kl_bwd ACCUMULATES into dzi / dzj (it composes with the contrastive backward) and overwrites dzo
        dzi, dzj, dzo = torch.zeros_like(xi), torch.zeros_like(xj), torch.empty_like(xo)
        stages.bwd(xi, xj, xo, norm, temperature, alpha, go, saved, dzi, dzj, dzo)
        return dzi.to(dti), dzj.to(dtj), dzo.to(dto), None, None, None, None, None


class DistributedRelicLoss(nn.Module):
    """RelicLoss (reference utils/losses.py:154-201, same ctor kwargs) over the global batch of a process group:
    contrastive term = DistributedSimclrLoss (rows all-gathered over NVLink, no gradient exchange), KL term = the
    batch-axis softmaxes over the all-gathered per-row logits (SURVEY.md §8e last row).  Every rank gets the gradient
    rows of its own inputs; the value equals the single-process RelicLoss on the rank-order concatenation."""

    def __init__(self, normalize=True, temperature=1.0, alpha=0.5, group=None, stages=None, kl_stages=None,
                 transport="auto"):
        super().__init__()
        self.normalize = normalize
        self.temperature = temperature
        self.alpha = alpha
        self.group = group
        self.contrastive = DistributedSimclrLoss(normalize, temperature, group=group, stages=stages, transport=transport)
        self.kl_stages = kl_stages if kl_stages is not None else RelicKlCudaStages()

    def forward(self, zi, zj, z_orig):
        kl = _RelicKlDistFn.apply(zi, zj, z_orig, self.normalize, self.temperature, self.alpha, self.group,
                                  self.kl_stages)
        return self.contrastive(zi, zj) + kl   # contrastive + alpha * KL (utils/losses.py:201)


# ======================================================================================================= DINO centre
@torch.no_grad()
def distributed_update_teacher_center(center, teacher_fvecs, momentum, group=None):
    """models/dino.py:136-141 over the global batch of a process group: the batch mean of the teacher outputs is taken
    over ALL ranks' rows (all-reduce of the K-vector of row sums + the row count, so unequal per-rank batches are
    weighted correctly), then the same EMA kernel as the single-GPU `update_teacher_center`.  `center` may be None on
    the first call, as in the reference.  Every rank ends up with the identical centre."""
    from .losses import update_teacher_center
    world, _ = _world_rank(group)
    local_mean = update_teacher_center(None, teacher_fvecs, 0.0)   # first call == plain batch mean (one kernel)
    if world > 1:
        rows = teacher_fvecs.numel() // teacher_fvecs.shape[-1]
        packed = torch.cat([local_mean * float(rows), local_mean.new_full((1,), float(rows))])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        local_mean = packed[:-1] / packed[-1]
    if center is None:
        return local_mean
    return update_teacher_center(center, local_mean.view(1, -1), momentum)


# ======================================================================================================= helpers
def _world_rank(group):
    if dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def _is_nccl(group):
    return dist.is_initialized() and dist.get_backend(group) == "nccl"


def _all_reduce_sum(t, group):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def _reduce_scatter_rows(full, rows, rank, group):
    """Sum `full` [world*rows x cols] over the ranks; return this rank's row slab [rows x cols]."""
    if _is_nccl(group):
        out = torch.empty(rows, full.shape[1], dtype=full.dtype, device=full.device)
        dist.reduce_scatter_tensor(out, full, op=dist.ReduceOp.SUM, group=group)
        return out
    dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)  # gloo (tests): no reduce-scatter
    return full[rank * rows:(rank + 1) * rows]


# ======================================================================================================= Barlow Twins
class BarlowCudaStages:
    """The product path of the distributed Barlow loss: the `ssvb_barlow_dist_*` stages of the C ABI."""

    def alloc_saved(self, n, d, dev):
        return C.byte_buffer(C.cached_size("ssvb_barlow_dist_saved_bytes", n, d), dev)

    def _ws(self, n, d, dev):
        nbytes = C.cached_size("ssvb_barlow_dist_workspace_bytes", n, d)
        return C.workspace("barlow_dist", nbytes, dev), nbytes

    def stats(self, zi, zj, normalize, stats_local, saved):
        n, d = zi.shape
        ws, nb = self._ws(n, d, zi.device)
        C.check(C.lib().ssvb_barlow_dist_stats(C.ptr(zi), C.ptr(zj), n, d, zi.stride(0), zj.stride(0), normalize,
                                               C.ptr(stats_local), C.ptr(saved), C.ptr(ws), nb,
                                               C.stream_ptr(zi.device)), "ssvb_barlow_dist_stats")

    def xcorr(self, zi, zj, normalize, stats_all, world, c_partial, saved):
        n, d = zi.shape
        C.check(C.lib().ssvb_barlow_dist_xcorr(C.ptr(zi), C.ptr(zj), n, d, zi.stride(0), zj.stride(0), normalize,
                                               C.ptr(stats_all), world, C.ptr(c_partial), C.ptr(saved),
                                               C.stream_ptr(zi.device)), "ssvb_barlow_dist_xcorr")

    def epilogue(self, c_rows, row0, lmbda, dc_rows, loss_partial, n):
        rows, d = c_rows.shape
        ws, nb = self._ws(n, d, c_rows.device)
        C.check(C.lib().ssvb_barlow_dist_epilogue(C.ptr(c_rows), row0, rows, d, lmbda, C.ptr(dc_rows),
                                                  C.ptr(loss_partial), C.ptr(ws), nb, C.stream_ptr(c_rows.device)),
                "ssvb_barlow_dist_epilogue")

    def bwd_gemm(self, zi, zj, n_global, normalize, dc, saved, colsum):
        n, d = zi.shape
        ws, nb = self._ws(n, d, zi.device)
        C.check(C.lib().ssvb_barlow_dist_bwd_gemm(C.ptr(zi), C.ptr(zj), n, n_global, d, zi.stride(0), zj.stride(0),
                                                  normalize, C.ptr(dc), C.ptr(saved), C.ptr(colsum), C.ptr(ws), nb,
                                                  C.stream_ptr(zi.device)), "ssvb_barlow_dist_bwd_gemm")

    def bwd_finish(self, zi, zj, n_global, normalize, colsum, grad_out, saved, dzi, dzj):
        n, d = zi.shape
        ws, nb = self._ws(n, d, zi.device)
        C.check(C.lib().ssvb_barlow_dist_bwd_finish(C.ptr(zi), C.ptr(zj), n, n_global, d, zi.stride(0), zj.stride(0),
                                                    normalize, C.ptr(colsum), C.ptr(grad_out), C.ptr(saved), C.ptr(dzi),
                                                    C.ptr(dzj), dzi.stride(0), dzj.stride(0), C.ptr(ws), nb,
                                                    C.stream_ptr(zi.device)), "ssvb_barlow_dist_bwd_finish")


    # ---- column-sharded variant (all-gather of the standardised rows; the D x D matrix never crosses NVLink)
    def standardize(self, zi, zj, normalize, stats_all, world, xi_slot, xj_slot, saved):
        n, d = zi.shape
        C.check(C.lib().ssvb_barlow_dist_standardize(C.ptr(zi), C.ptr(zj), n, d, zi.stride(0), zj.stride(0), normalize,
                                                     C.ptr(stats_all), world, C.ptr(xi_slot), C.ptr(xj_slot),
                                                     C.ptr(saved), C.stream_ptr(zi.device)), "ssvb_barlow_dist_standardize")

    def _cs_ws(self, ng, d, ncols, dev):
        nbytes = C.cached_size("ssvb_barlow_cs_workspace_bytes", ng, d, ncols)
        return C.workspace("barlow_cs", nbytes, dev), nbytes

    def cs_fwd(self, xa_all, xb_all, col0, ncols, lmbda, dc_slab, loss_partial):
        ng, d = xa_all.shape
        ws, nb = self._cs_ws(ng, d, ncols, xa_all.device)
        C.check(C.lib().ssvb_barlow_cs_fwd(C.ptr(xa_all), C.ptr(xb_all), ng, d, col0, ncols, lmbda, C.ptr(dc_slab),
                                           C.ptr(loss_partial), C.ptr(ws), nb, C.stream_ptr(xa_all.device)),
                "ssvb_barlow_cs_fwd")

    def cs_bwd(self, xa_all, xb_all, dc_slab, col0, ncols, saved, n_local, view_b, grad_out, dxb_slab):
        ng, d = xa_all.shape
        ws, nb = self._cs_ws(ng, d, ncols, xa_all.device)
        C.check(C.lib().ssvb_barlow_cs_bwd(C.ptr(xa_all), C.ptr(xb_all), C.ptr(dc_slab), ng, d, col0, ncols, C.ptr(saved),
                                           n_local, view_b, C.ptr(grad_out), C.ptr(dxb_slab), C.ptr(ws), nb,
                                           C.stream_ptr(xa_all.device)), "ssvb_barlow_cs_bwd")

    def cs_finish(self, recv, world, n_local, ncols, x, normalize, saved, view, dx):
        C.check(C.lib().ssvb_barlow_cs_finish(C.ptr(recv), world, n_local, ncols, C.ptr(x), x.stride(0), normalize,
                                              C.ptr(saved), view, C.ptr(dx), dx.stride(0), C.stream_ptr(x.device)),
                "ssvb_barlow_cs_finish")


def _all_to_all_blocks(send, group, nccl):
    """send: [world][block...] -> recv[q] = the block rank q addressed to this rank."""
    recv = torch.empty_like(send)
    if nccl:
        dist.all_to_all_single(recv, send, group=group)
    else:  # gloo (tests): emulate with an all-gather
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send.contiguous(), group=group)
        for q in range(world):
            recv[q].copy_(parts[q][rank])
    return recv


class _BarlowColShardFn(torch.autograd.Function):
    """Column-sharded global-batch Barlow Twins (see DistributedBarlowLoss, mode="colshard")."""

    @staticmethod
    def forward(ctx, zi, zj, normalize, lmbda, group, stages):
        world, rank = _world_rank(group)
        cuda = isinstance(stages, BarlowCudaStages)
        if cuda:
            C.require_cuda(zi, zj)
            xi, xj = C.as_f32_rows(zi), C.as_f32_rows(zj)
        else:
            xi, xj = zi.detach().float().contiguous(), zj.detach().float().contiguous()
        if xi.shape != xj.shape or xi.dim() != 2:
            raise ValueError("BarlowLoss expects two [N, D] tensors of the same shape")
        n, d = xi.shape
        if d % world:
            raise ValueError("colshard mode needs D divisible by the world size")
        ng, ncols = n * world, d // world
        col0 = rank * ncols
        dev = xi.device
        norm = int(bool(normalize))
        saved = stages.alloc_saved(n, d, dev)
        stats_all = torch.empty(world, 2, 2, d, dtype=torch.float32, device=dev)
        stages.stats(xi, xj, norm, stats_all[rank], saved)
        if world > 1:
            _gather_slots(stats_all.view(world, 4 * d), stats_all[rank].view(1, 4 * d), group, inplace=cuda)
        # standardised rows straight into this rank's slot of the gathered matrices, then ONE all-gather per view
        xt = torch.empty(2, ng, d, dtype=torch.bfloat16, device=dev)
        my = slice(rank * n, (rank + 1) * n)
        stages.standardize(xi, xj, norm, stats_all, world, xt[0, my], xt[1, my], saved)
        if world > 1:
            _gather_slots(xt[0], xt[0, my], group, inplace=cuda)
            _gather_slots(xt[1], xt[1, my], group, inplace=cuda)
        # column slab of C (loss + dC) and of C^T (dC^T): both gradient slabs will be complete, no reduction of C at all
        dc = torch.empty(2, d, ncols, dtype=torch.bfloat16, device=dev)
        parts = torch.zeros(world, dtype=torch.float32, device=dev)
        stages.cs_fwd(xt[0], xt[1], col0, ncols, float(lmbda), dc[0], parts[rank:rank + 1])
        stages.cs_fwd(xt[1], xt[0], col0, ncols, float(lmbda), dc[1], None)
        if world > 1:
            _gather_slots(parts.view(world, 1), parts[rank:rank + 1].view(1, 1), group, inplace=cuda)
        loss = parts.sum()  # same values, same order on every rank
        ctx.save_for_backward(xi, xj, saved, xt, dc)
        ctx.cfg = (norm, world, rank, group, stages, cuda, zi.dtype, zj.dtype)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        xi, xj, saved, xt, dc = ctx.saved_tensors
        norm, world, rank, group, stages, cuda, dti, dtj = ctx.cfg
        n, d = xi.shape
        ng, ncols = n * world, d // world
        col0 = rank * ncols
        go = C.f32_scalar(grad_out)
        slabs = torch.empty(2, ng, ncols, dtype=torch.float32, device=xi.device)
        # d Xj~[:, slab] = Xi~ dC[:, slab];  d Xi~[:, slab] = Xj~ dC^T[:, slab]; standardisation backward is column-local
        stages.cs_bwd(xt[0], xt[1], dc[0], col0, ncols, saved, n, 1, go, slabs[1])
        stages.cs_bwd(xt[1], xt[0], dc[1], col0, ncols, saved, n, 0, go, slabs[0])
        dzi, dzj = torch.empty_like(xi), torch.empty_like(xj)
        for view, (x, out) in enumerate(((xi, dzi), (xj, dzj))):
            send = slabs[view].view(world, n, ncols)  # row block q of the slab belongs to rank q
            recv = _all_to_all_blocks(send, group, cuda) if world > 1 else send
            stages.cs_finish(recv, world, n, ncols, x, norm, saved, view, out)
        return dzi.to(dti), dzj.to(dtj), None, None, None, None


class _BarlowDistFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, zi, zj, normalize, lmbda, group, stages):
        world, rank = _world_rank(group)
        cuda = isinstance(stages, BarlowCudaStages)
        if cuda:
            C.require_cuda(zi, zj)
            xi, xj = C.as_f32_rows(zi), C.as_f32_rows(zj)
        else:
            xi, xj = zi.detach().float().contiguous(), zj.detach().float().contiguous()
        if xi.shape != xj.shape or xi.dim() != 2:
            raise ValueError("BarlowLoss expects two [N, D] tensors of the same shape")
        n, d = xi.shape
        dev = xi.device
        norm = int(bool(normalize))
        saved = stages.alloc_saved(n, d, dev)
        # 1. local column moments -> all-gather (tiny) -> every rank combines them in rank order
        stats_all = torch.empty(world, 2, 2, d, dtype=torch.float32, device=dev)
        stages.stats(xi, xj, norm, stats_all[rank], saved)
        if world > 1:
            _gather_slots(stats_all.view(world, 4 * d), stats_all[rank].view(1, 4 * d), group, inplace=cuda)
        # 2. partial cross-correlation of the local rows, summed over ranks: reduce-scatter -> this rank's row slab
        c_partial = torch.empty(d, d, dtype=torch.float32, device=dev)
        stages.xcorr(xi, xj, norm, stats_all, world, c_partial, saved)
        sharded = world > 1 and d % world == 0
        dc = torch.empty(d, d, dtype=torch.bfloat16, device=dev)
        parts = torch.zeros(world if sharded else 1, dtype=torch.float32, device=dev)
        if sharded:
            rows = d // world
            c_rows = _reduce_scatter_rows(c_partial, rows, rank, group)
            # 3. loss terms + dC of the slab, then all-gather dC in bf16 (second half of the all-reduce at half width)
            stages.epilogue(c_rows, rank * rows, float(lmbda), dc[rank * rows:(rank + 1) * rows], parts[rank:rank + 1], n)
            _gather_slots(dc, dc[rank * rows:(rank + 1) * rows], group, inplace=cuda)
            _gather_slots(parts.view(world, 1), parts[rank:rank + 1].view(1, 1), group, inplace=cuda)
            loss = parts.sum()  # same values, same order on every rank -> identical global loss, no all-reduce
        else:
            if world > 1:
                _all_reduce_sum(c_partial, group)
            stages.epilogue(c_partial, 0, float(lmbda), dc, parts, n)
            loss = parts[0].clone()
        del c_partial
        ctx.save_for_backward(xi, xj, saved, dc)
        ctx.cfg = (norm, world, group, stages, zi.dtype, zj.dtype)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        xi, xj, saved, dc = ctx.saved_tensors
        norm, world, group, stages, dti, dtj = ctx.cfg
        n, d = xi.shape
        go = C.f32_scalar(grad_out)
        colsum = torch.empty(2, 2, d, dtype=torch.float32, device=xi.device)
        stages.bwd_gemm(xi, xj, n * world, norm, dc, saved, colsum)
        _all_reduce_sum(colsum, group)
        dzi, dzj = torch.empty_like(xi), torch.empty_like(xj)
        stages.bwd_finish(xi, xj, n * world, norm, colsum, go, saved, dzi, dzj)
        return dzi.to(dti), dzj.to(dtj), None, None, None, None


class DistributedBarlowLoss(nn.Module):
    """BarlowLoss (reference utils/losses.py:120-142, same ctor kwargs) over the global batch of a process group:
    batch rows sharded, column statistics all-gathered, the D x D cross-correlation all-reduced as
    reduce-scatter(fp32) -> fused loss/dC epilogue on the slab -> all-gather(bf16), the two column reductions of the
    standardisation backward all-reduced.  Every rank gets the gradient rows of its own inputs (no 1/world rescale)."""

    def __init__(self, normalize=True, off_diagonal_weight=0.005, group=None, stages=None, mode="allreduce"):
        """mode "allreduce": the all-reduce of the cross-correlation described above (BASELINE.json's wording);
        mode "colshard": all-gather the standardised bf16 rows instead and give every rank a column slab of C and C^T
        (~4x less NVLink traffic at D = 8192, no all-reduce of C, gradients re-sharded by one small all-to-all);
        needs D % (8 * world) == 0."""
        super().__init__()
        if mode not in ("allreduce", "colshard"):
            raise ValueError("mode must be 'allreduce' or 'colshard'")
        self.normalize = normalize
        self.lmbda = off_diagonal_weight
        self.group = group
        self.stages = stages if stages is not None else BarlowCudaStages()
        self.mode = mode

    def forward(self, z_i, z_j):
        fn = _BarlowColShardFn if self.mode == "colshard" else _BarlowDistFn
        return fn.apply(z_i, z_j, self.normalize, self.lmbda, self.group, self.stages)


# ======================================================================================================= SwAV / Sinkhorn
class SwavCudaStages:
    """The product path of the distributed SwAV loss / Sinkhorn: C-ABI stages."""

    def kpad(self, k):
        return C.cached_size("ssvb_swav_kpad", k)

    def alloc_saved(self, nb, nbank, k, d, dev):
        return C.byte_buffer(C.cached_size("ssvb_swav_saved_bytes", nb, nbank, k, d), dev)

    def sk_pass(self, phase, scores, b_global, k, eps, alpha, smax, u_local, codes):
        dev = scores.device
        nbytes = C.cached_size("ssvb_sinkhorn_workspace_bytes", scores.shape[0], k)
        ws = C.workspace("sinkhorn_dist", nbytes, dev)
        C.check(C.lib().ssvb_sinkhorn_dist_pass(phase, C.ptr(scores), scores.shape[0], b_global, k, scores.stride(0),
                                                eps, C.ptr(alpha), C.ptr(smax), C.ptr(u_local), C.ptr(codes),
                                                codes.stride(0) if codes is not None else 0, C.ptr(ws), nbytes,
                                                C.stream_ptr(dev)), "ssvb_sinkhorn_dist_pass")

    def sk_alpha(self, u_all_view, world, rank_stride, k, phase0, eps, alpha, smax):
        C.check(C.lib().ssvb_sinkhorn_dist_alpha(C.ptr(u_all_view), world, rank_stride, k, int(phase0), eps,
                                                 C.ptr(alpha), C.ptr(smax), C.stream_ptr(alpha.device)),
                "ssvb_sinkhorn_dist_alpha")

    def scores(self, z1, z2, bank, protos, scores, saved):
        nb, d = z1.shape
        nbank = bank.shape[0] if bank is not None else 0
        C.check(C.lib().ssvb_swav_dist_scores(C.ptr(z1), C.ptr(z2), C.ptr(bank), C.ptr(protos), nb, nbank,
                                              protos.shape[0], d, z1.stride(0), z2.stride(0),
                                              bank.stride(0) if bank is not None else 0, protos.stride(0),
                                              C.ptr(scores), C.ptr(saved), C.stream_ptr(z1.device)),
                "ssvb_swav_dist_scores")

    def ce(self, scores, codes, nb, nbank, bp_global, k, d, temperature, loss_local, saved):
        dev = scores.device
        nbytes = C.cached_size("ssvb_swav_workspace_bytes", nb, nbank, k, d)
        ws = C.workspace("swav", nbytes, dev)
        C.check(C.lib().ssvb_swav_dist_ce(C.ptr(scores), C.ptr(codes), nb, nbank, bp_global, k, d, temperature,
                                          C.ptr(loss_local), C.ptr(saved), C.ptr(ws), nbytes, C.stream_ptr(dev)),
                "ssvb_swav_dist_ce")

    def bwd(self, z1, z2, bank, protos, temperature, grad_out, saved, dz1, dz2, dproto):
        nb, d = z1.shape
        nbank = bank.shape[0] if bank is not None else 0
        k = protos.shape[0]
        dev = z1.device
        nbytes = C.cached_size("ssvb_swav_workspace_bytes", nb, nbank, k, d)
        ws = C.workspace("swav", nbytes, dev)
        C.check(C.lib().ssvb_swav_bwd(C.ptr(z1), C.ptr(z2), C.ptr(bank), C.ptr(protos), nb, nbank, k, d, z1.stride(0),
                                      z2.stride(0), bank.stride(0) if bank is not None else 0, protos.stride(0),
                                      temperature, C.ptr(grad_out), C.ptr(saved), C.ptr(dz1), C.ptr(dz2), C.ptr(dproto),
                                      dz1.stride(0), dz2.stride(0), dproto.stride(0), C.ptr(ws), nbytes,
                                      C.stream_ptr(dev)), "ssvb_swav_bwd")


def _dist_sinkhorn(stages, scores_views, codes_views, k, eps, n_iters, group, inplace):
    """Row-sharded Sinkhorn-Knopp on one or more views at once (their marginals travel in ONE all-gather per pass).
    scores_views / codes_views: lists of [b_local, >=k] fp32 tensors (same b_local on every rank)."""
    world, rank = _world_rank(group)
    nv = len(scores_views)
    b_local = scores_views[0].shape[0]
    b_global = b_local * world
    dev = scores_views[0].device
    u_all = torch.empty(world, nv, k + 1, dtype=torch.float32, device=dev)
    alpha = torch.empty(nv, k, dtype=torch.float32, device=dev)
    smax = torch.empty(nv, 1, dtype=torch.float32, device=dev)
    eps = float(eps)

    def exchange(phase0):
        if world > 1:
            _gather_slots(u_all.view(world, nv * (k + 1)), u_all[rank].view(1, nv * (k + 1)), group, inplace=inplace)
        for v in range(nv):
            stages.sk_alpha(u_all[0, v], world, nv * (k + 1), k, phase0, eps, alpha[v], smax[v])

    for v in range(nv):
        stages.sk_pass(0, scores_views[v], b_global, k, eps, None, None, u_all[rank, v], codes_views[v])
    exchange(True)
    if n_iters <= 0:
        alpha.fill_(1.0)  # no iterations: codes = E / rowsum(E)
    for _ in range(1, int(n_iters)):
        for v in range(nv):
            stages.sk_pass(1, scores_views[v], b_global, k, eps, alpha[v], smax[v], u_all[rank, v], codes_views[v])
        exchange(False)
    for v in range(nv):
        stages.sk_pass(2, scores_views[v], b_global, k, eps, alpha[v], smax[v], None, codes_views[v])


class _SwavDistFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, prototypes, bank, temperature, eps, n_iters, group, stages, reduce_dproto=False):
        ctx.reduce_dproto = bool(reduce_dproto)
        world, rank = _world_rank(group)
        cuda = isinstance(stages, SwavCudaStages)
        if cuda:
            C.require_cuda(z1, z2, prototypes, bank)
            x1, x2, pc = C.as_f32_rows(z1), C.as_f32_rows(z2), C.as_f32_rows(prototypes)
            bk = C.as_f32_rows(bank.detach()) if bank is not None and bank.shape[0] > 0 else None
        else:
            x1, x2, pc = (t.detach().float().contiguous() for t in (z1, z2, prototypes))
            bk = bank.detach().float().contiguous() if bank is not None and bank.shape[0] > 0 else None
        if x1.shape != x2.shape or x1.dim() != 2 or pc.dim() != 2 or pc.shape[1] != x1.shape[1]:
            raise ValueError("SwavLoss expects z_1/z_2 [B, d] and prototypes [K, d]")
        nb, d = x1.shape
        nbank = bk.shape[0] if bk is not None else 0
        k = pc.shape[0]
        bp = nb + nbank
        dev = x1.device
        kp = stages.kpad(k)
        saved = stages.alloc_saved(nb, nbank, k, d, dev)
        scores = torch.empty(2 * bp, kp, dtype=torch.float32, device=dev)
        codes = torch.empty(2 * bp, kp, dtype=torch.float32, device=dev)
        stages.scores(x1, x2, bk, pc, scores, saved)
        _dist_sinkhorn(stages, [scores[:bp], scores[bp:]], [codes[:bp], codes[bp:]], k, eps, n_iters, group, cuda)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        stages.ce(scores, codes, nb, nbank, bp * world, k, d, float(temperature), loss, saved)
        _all_reduce_sum(loss, group)
        ctx.save_for_backward(x1, x2, pc, saved, *([bk] if bk is not None else []))
        ctx.cfg = (float(temperature), group, stages, z1.dtype, z2.dtype, prototypes.dtype)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        x1, x2, pc, saved, *rest = ctx.saved_tensors
        bk = rest[0] if rest else None
        temperature, group, stages, dt1, dt2, dtp = ctx.cfg
        go = C.f32_scalar(grad_out)
        dz1, dz2, dpc = torch.empty_like(x1), torch.empty_like(x2), torch.empty_like(pc)
        stages.bwd(x1, x2, bk, pc, temperature, go, saved, dz1, dz2, dpc)
        # Prototypes are replicated.  By default this rank returns its LOCAL contribution (the rows it owns), exactly like
        # dz1 / dz2 feed only local contributions into the encoder's parameter gradients: whatever uniform cross-rank
        # reduction the training loop applies to its replicated parameters (DDP mean, a SUM all-reduce) then treats
        # encoder and prototypes alike.  reduce_prototype_grad=True instead returns the complete global gradient on
        # every rank - for loops that EXCLUDE the prototypes from gradient synchronisation.
        if ctx.reduce_dproto:
            _all_reduce_sum(dpc, group)
        return dz1.to(dt1), dz2.to(dt2), dpc.to(dtp), None, None, None, None, None, None, None


class DistributedSwavLoss(nn.Module):
    """SwavLoss (reference utils/losses.py:204-235, same ctor kwargs) over the global batch of a process group: sample
    rows (and each rank's bank rows) sharded, prototypes replicated.  Per Sinkhorn pass ONE all-gather of the K-vector
    prototype marginals of both views (combined in rank order = their all-reduce); all-reduce of the scalar loss.
    Every rank gets the gradient rows of its own inputs and its local contribution to the prototype gradient (see
    `reduce_prototype_grad`)."""

    def __init__(self, temperature=0.1, sinkhorn_eps=0.05, sinkhorn_iters=3, group=None, stages=None,
                 reduce_prototype_grad=False):
        """reduce_prototype_grad=False (default): the prototype gradient returned on each rank is that rank's LOCAL
        contribution, like the row gradients - reduce it together with the encoder's parameter gradients (DDP or an
        all-reduce).  True: every rank gets the complete global prototype gradient (then keep the prototypes OUT of the
        gradient synchronisation, or they end up `world` times too large relative to the encoder)."""
        super().__init__()
        self.reduce_prototype_grad = reduce_prototype_grad
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.temperature = temperature
        self.n_iters = sinkhorn_iters
        self.eps = sinkhorn_eps
        self.group = group
        self.stages = stages if stages is not None else SwavCudaStages()

    @torch.no_grad()
    def compute_codes_sinkhorn(self, scores):
        """Codes of this rank's rows of the row-sharded global score matrix (utils/losses.py:213-224 on the concatenation)."""
        cuda = isinstance(self.stages, SwavCudaStages)
        if cuda:
            C.require_cuda(scores)
        s = scores.detach()
        if s.dtype != torch.float32 or s.stride(1) != 1:
            s = s.float().contiguous()
        codes = torch.empty(s.shape[0], s.shape[1], dtype=torch.float32, device=s.device)
        _dist_sinkhorn(self.stages, [s], [codes], s.shape[1], self.eps, self.n_iters, self.group, cuda)
        return codes

    def forward(self, z_1, z_2, prototypes, bank_features=None):
        return _SwavDistFn.apply(z_1, z_2, prototypes, bank_features, self.temperature, self.eps, self.n_iters,
                                 self.group, self.stages, self.reduce_prototype_grad)


# ======================================================================================================= MoCo (sharded queue)
class MocoCudaStages:
    """The product path of the sharded-queue MoCo loss: the `ssvb_moco_dist_*` stages of the C ABI."""

    def dpad(self, d):
        return C.cached_size("ssvb_ntxent_dpad", d)

    def npad(self, n_global):
        return C.cached_size("ssvb_moco_dist_npad", n_global)

    def _ws(self, n_global, k_local, d, dev):
        nbytes = C.cached_size("ssvb_moco_dist_workspace_bytes", n_global, k_local, d)
        return C.workspace("moco_dist", nbytes, dev), nbytes

    def prep(self, q, k, normalize, world, rank, qhat_all, rowstat):
        n, d = q.shape
        C.check(C.lib().ssvb_moco_dist_prep(C.ptr(q), C.ptr(k), n, d, q.stride(0), k.stride(0), normalize, world, rank,
                                            C.ptr(qhat_all), C.ptr(rowstat), C.stream_ptr(q.device)),
                "ssvb_moco_dist_prep")

    def shard_fwd(self, qhat_all, n_global, shard, shadow, d, temperature, rowstat, n_local, part_local):
        ws, nb = self._ws(n_global, shard.shape[0], d, shard.device)
        C.check(C.lib().ssvb_moco_dist_shard_fwd(C.ptr(qhat_all), n_global, C.ptr(shard), C.ptr(shadow), shard.shape[0],
                                                 d, shard.stride(0), temperature, C.ptr(rowstat), n_local,
                                                 C.ptr(part_local), C.ptr(ws), nb, C.stream_ptr(shard.device)),
                "ssvb_moco_dist_shard_fwd")

    def finalize(self, part_all, world, n_local, temperature, lse2_all, loss, k_local, d):
        ws, nb = self._ws(world * n_local, k_local, d, part_all.device)
        C.check(C.lib().ssvb_moco_dist_finalize(C.ptr(part_all), world, n_local, temperature, C.ptr(lse2_all),
                                                C.ptr(loss), C.ptr(ws), nb, C.stream_ptr(part_all.device)),
                "ssvb_moco_dist_finalize")

    def shard_bwd(self, qhat_all, n_global, shard, shadow, d, temperature, lse2_all, dacc_partial):
        ws, nb = self._ws(n_global, shard.shape[0], d, shard.device)
        C.check(C.lib().ssvb_moco_dist_shard_bwd(C.ptr(qhat_all), n_global, C.ptr(shard), C.ptr(shadow), shard.shape[0],
                                                 d, shard.stride(0), temperature, C.ptr(lse2_all), C.ptr(dacc_partial),
                                                 C.ptr(ws), nb, C.stream_ptr(shard.device)), "ssvb_moco_dist_shard_bwd")

    def finish(self, q, k, n_global, normalize, temperature, rowstat, lse2_local, dacc_local, grad_out, dq, dk):
        n, d = q.shape
        C.check(C.lib().ssvb_moco_dist_finish(C.ptr(q), C.ptr(k), n, n_global, d, q.stride(0), k.stride(0), normalize,
                                              temperature, C.ptr(rowstat), C.ptr(lse2_local), C.ptr(dacc_local),
                                              C.ptr(grad_out), C.ptr(dq), C.ptr(dk), dq.stride(0), dk.stride(0),
                                              C.stream_ptr(q.device)), "ssvb_moco_dist_finish")


class _MocoDistFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query, keys, shard, normalize, temperature, group, stages):
        world, rank = _world_rank(group)
        cuda = isinstance(stages, MocoCudaStages)
        if cuda:
            C.require_cuda(query, keys, shard)
            q, k = C.as_f32_rows(query), C.as_f32_rows(keys)
            mem = C.as_f32_rows(shard.detach())
            from .banks import lookup_shadow
            shadow = lookup_shadow(shard) if mem.data_ptr() == shard.data_ptr() else None
        else:
            q, k, mem = (t.detach().float().contiguous() for t in (query, keys, shard))
            shadow = None
        if q.shape != k.shape or q.dim() != 2 or mem.dim() != 2 or mem.shape[1] != q.shape[1]:
            raise ValueError("MocoLoss expects query/keys [N, d] and a queue shard [K_local, d]")
        n, d = q.shape
        ng = n * world
        dev = q.device
        norm = int(bool(normalize))
        tau = float(temperature)
        npad, dpad = stages.npad(ng), stages.dpad(d)
        qhat_all = torch.empty(npad, dpad, dtype=torch.bfloat16, device=dev)
        rowstat = torch.empty(3, n, dtype=torch.float32, device=dev)
        stages.prep(q, k, norm, world, rank, qhat_all, rowstat)
        if world > 1:
            _gather_slots(qhat_all[:ng], qhat_all[rank * n:(rank + 1) * n], group, inplace=cuda)
        blk = 2 * ng + n
        part_all = torch.empty(world, blk, dtype=torch.float32, device=dev)
        stages.shard_fwd(qhat_all, ng, mem, shadow, d, tau, rowstat, n, part_all[rank])
        if world > 1:
            _gather_slots(part_all, part_all[rank:rank + 1], group, inplace=cuda)
        lse2_all = torch.zeros(npad, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        stages.finalize(part_all, world, n, tau, lse2_all, loss, mem.shape[0], d)
        ctx.save_for_backward(q, k, mem, qhat_all, rowstat, lse2_all)
        ctx.shadow = shadow
        ctx.cfg = (norm, tau, world, rank, group, stages, query.dtype, keys.dtype)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        q, k, mem, qhat_all, rowstat, lse2_all = ctx.saved_tensors
        norm, tau, world, rank, group, stages, dtq, dtk = ctx.cfg
        n, d = q.shape
        ng = n * world
        dev = q.device
        go = C.f32_scalar(grad_out)
        dacc = torch.empty(qhat_all.shape[0], qhat_all.shape[1], dtype=torch.float32, device=dev)
        stages.shard_bwd(qhat_all, ng, mem, ctx.shadow, d, tau, lse2_all, dacc)
        # sum_j p_aj m_j over ALL shards, for this rank's own query rows only: reduce-scatter
        dacc_local = _reduce_scatter_rows(dacc[:ng], n, rank, group) if world > 1 else dacc[:n]
        dq, dk = torch.empty_like(q), torch.empty_like(k)
        stages.finish(q, k, ng, norm, tau, rowstat, lse2_all[rank * n:(rank + 1) * n], dacc_local, go, dq, dk)
        return dq.to(dtq), dk.to(dtk), None, None, None, None, None


class DistributedMocoLoss(nn.Module):
    """MocoLoss (reference utils/losses.py:49-72, same ctor kwargs) with the queue SHARDED over the ranks of a process
    group: forward(query, keys, queue_shard) where queue_shard is this rank's K/world rows (ShardedMemoryBank).
    all-gather of the bf16 queries, all-gather of the per-shard (max, sum-exp) partials (combined in rank order: the
    global loss is identical on every rank), reduce-scatter of the query-gradient partials."""

    def __init__(self, normalize=True, temperature=1.0, group=None, stages=None):
        super().__init__()
        self.normalize = normalize
        self.temperature = temperature
        self.group = group
        self.stages = stages if stages is not None else MocoCudaStages()

    def forward(self, query, keys, memory_vectors):
        return _MocoDistFn.apply(query, keys, memory_vectors, self.normalize, self.temperature, self.group, self.stages)


class ShardedMemoryBank:
    """MemoryBank (reference models/moco.py:23-39) range-partitioned over the ranks of a process group: rank s holds
    global rows [s*K/world, (s+1)*K/world).  `add_batch(keys)` all-gathers the local keys (rank order = the
    concatenated global batch) and every rank writes the rows of [ptr, ptr + N_global) mod K that fall into its shard;
    `ptr` is the single-process ring's pointer, identical on every rank (bit-exact bookkeeping)."""

    def __init__(self, queue_size, feature_size, group=None, device=None):
        from .banks import _Ring
        world, rank = _world_rank(group)
        if queue_size % world:
            raise ValueError("queue_size must be divisible by the world size")
        self.size = queue_size
        self.group = group
        self.world, self.rank = world, rank
        self.shard_rows = queue_size // world
        self.shard_lo = rank * self.shard_rows
        self._ring = _Ring()
        self._ring._init_ring(self.shard_rows, feature_size, device, normalize=True, with_shadow=True)
        self.ptr = 0

    @property
    def bank(self):
        return self._ring._data

    def get_vectors(self):
        return self._ring._data

    def add_batch(self, batch):
        import ctypes
        r = self._ring
        b = C.as_f32_rows(batch.detach().to(r._device, non_blocking=True))
        if self.world > 1:
            allb = torch.empty(self.world * b.shape[0], b.shape[1], dtype=torch.float32, device=r._device)
            dist.all_gather_into_tensor(allb, b, group=self.group)
        else:
            allb = b
        rows, dim = r._data.shape
        new_ptr = ctypes.c_int64(-1)
        shadow_ok = r._shadow is not None and r._shadow_version == r._data._version
        with torch.cuda.device(r._device):
            C.check(C.lib().ssvb_ring_enqueue_shard(C.ptr(r._data), C.ptr(r._shadow) if shadow_ok else None, self.size,
                                                    self.shard_lo, rows, dim, r._data.stride(0), C.ptr(allb),
                                                    allb.shape[0], allb.stride(0), int(self.ptr), 1,
                                                    ctypes.cast(ctypes.pointer(new_ptr), ctypes.c_void_p),
                                                    C.stream_ptr(r._device)), "ssvb_ring_enqueue_shard")
        self.ptr = int(new_ptr.value)
