"""SURVEY.md §8(f) rank 1 - the step on either side of the loss, fused.

The reference's BYOL and SimSiam heads end in `F.normalize` (models/byol.py:47,59; models/simsiam.py:48,69) and the
loss then re-reads the unit rows.  `NormalizedMSELoss` / `NormalizedSimSiamLoss` take the RAW head outputs and run
normalise + loss in one pass forward and one pass backward (csrc/rowwise.cu `rowdot_norm_*`): value and gradients equal
`MSELoss()(F.normalize(o), F.normalize(t))` / `SimSiamLoss()(F.normalize(o), F.normalize(t))`.  To use them, drop the
trailing `F.normalize` from the head (or feed the pre-normalisation activations).

`graphed(module, *sample_inputs)` captures a loss module's forward AND backward in CUDA graphs (static shapes): the
small-batch configurations the reference actually trains at (bs 256-512, configs/*.yaml) are launch-bound, a graph
replay removes the per-launch Python / driver cost.
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import _cabi as C


class _RowdotNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, o, t, kind, norm_o, norm_t):
        C.require_cuda(o, t)
        if o.shape != t.shape or o.dim() != 2:
            raise ValueError("the fused normalised losses expect two [N, d] tensors of the same shape")
        oo, tt = C.as_f32_rows(o), C.as_f32_rows(t)
        n, d = oo.shape
        L = C.lib()
        dev = oo.device
        with C.on_device(dev):
            saved = C.byte_buffer(C.cached_size("ssvb_rowdot_norm_saved_bytes", n), dev)
            ws_bytes = C.cached_size("ssvb_rowdot_workspace_bytes", n, d)
            ws = C.workspace("rowdot", ws_bytes, dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            C.check(L.ssvb_rowdot_norm_fwd(kind, C.ptr(oo), C.ptr(tt), n, d, oo.stride(0), tt.stride(0), int(norm_o),
                                           int(norm_t), C.ptr(loss), C.ptr(saved), C.ptr(ws), ws_bytes,
                                           C.stream_ptr(dev)), "ssvb_rowdot_norm_fwd")
        ctx.save_for_backward(oo, tt, saved)
        ctx.cfg = (kind, int(norm_o), int(norm_t), o.dtype, t.dtype)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        oo, tt, saved = ctx.saved_tensors
        kind, norm_o, norm_t, dto, dtt = ctx.cfg
        need_o, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_o or need_t):
            return None, None, None, None, None
        n, d = oo.shape
        dev = oo.device
        with C.on_device(dev):
            go = C.f32_scalar(grad_out)
            d_o = torch.empty_like(oo) if need_o else None
            d_t = torch.empty_like(tt) if need_t else None
            C.check(C.lib().ssvb_rowdot_norm_bwd(kind, C.ptr(oo), C.ptr(tt), n, d, oo.stride(0), tt.stride(0), norm_o,
                                                 norm_t, C.ptr(go), C.ptr(saved), C.ptr(d_o), C.ptr(d_t),
                                                 d_o.stride(0) if need_o else 0, d_t.stride(0) if need_t else 0,
                                                 C.stream_ptr(dev)), "ssvb_rowdot_norm_bwd")
        return (d_o.to(dto) if need_o else None, d_t.to(dtt) if need_t else None, None, None, None)


class NormalizedMSELoss(nn.Module):
    """BYOL: nn.MSELoss()(F.normalize(input), F.normalize(target)) on the raw head outputs (models/byol.py:47,59,89,
    129-130), normalisation fused into the loss kernels."""

    def __init__(self, normalize_input=True, normalize_target=True):
        super().__init__()
        self.normalize_input, self.normalize_target = normalize_input, normalize_target

    def forward(self, input, target):  # noqa: A002
        return _RowdotNormFn.apply(input, target, 0, self.normalize_input, self.normalize_target)


class NormalizedSimSiamLoss(nn.Module):
    """SimSiam: SimSiamLoss()(F.normalize(online), F.normalize(target)) on the raw head outputs (models/simsiam.py:48,69;
    utils/losses.py:150-151), normalisation fused into the loss kernels."""

    def __init__(self, normalize_online=True, normalize_target=True):
        super().__init__()
        self.normalize_online, self.normalize_target = normalize_online, normalize_target

    def forward(self, online_output, target_output):
        return _RowdotNormFn.apply(online_output, target_output, 1, self.normalize_online, self.normalize_target)


def graphed(module, *sample_inputs, num_warmup_iters=3):
    """CUDA-graphed version of a loss module for fixed input shapes: forward and backward are captured once and replayed
    (torch.cuda.make_graphed_callables; every ssv_b200 op is capture-safe: no host sync, no allocation outside torch's
    caching allocator, workspaces are persistent).  `sample_inputs` fix shapes / dtypes / requires_grad; the returned
    callable is used exactly like the module.  Stateful side inputs that change between steps (a bank that is enqueued)
    must be passed as arguments, not captured."""
    for t in sample_inputs:
        if torch.is_tensor(t):
            C.require_cuda(t)
    return torch.cuda.make_graphed_callables(module, tuple(sample_inputs), num_warmup_iters=num_warmup_iters)


class SelaLabeler:
    """SeLA self-labelling state + step (reference models/sela.py:72-73 state, :146-166 step): `alpha` [K, 1] and `beta`
    [B, 1] start as N(0, 1) draws like the reference's and are carried from batch to batch; `step(logits)` runs
    P = pow(log_softmax(logits), lambda)^T, `num_iters` alternating reciprocal-matvec updates and the final argmax in
    ONE kernel launch and returns the pseudo-labels of the batch (int64 [B])."""

    def __init__(self, num_clusters, batch_size, lmbd, device=None):
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.lmbd = float(lmbd)
        self.alpha = torch.empty(num_clusters, 1, dtype=torch.float32).normal_(0, 1).to(dev)   # sela.py:72
        self.beta = torch.empty(batch_size, 1, dtype=torch.float32).normal_(0, 1).to(dev)      # sela.py:73

    @torch.no_grad()
    def step(self, logits, num_iters=80):
        C.require_cuda(logits, self.alpha, self.beta)
        x = C.as_f32_rows(logits.detach())
        b, k = x.shape
        if self.alpha.numel() != k or self.beta.numel() != b:
            raise ValueError(f"SelaLabeler was built for [{self.beta.numel()} x {self.alpha.numel()}] logits, got {tuple(x.shape)}")
        dev = x.device
        labels = torch.empty(b, dtype=torch.int64, device=dev)
        with C.on_device(dev):
            ws_bytes = C.cached_size("ssvb_sela_workspace_bytes", b, k)
            ws = C.workspace("sela", ws_bytes, dev)
            C.check(C.lib().ssvb_sela_self_label(C.ptr(x), b, k, x.stride(0), self.lmbd, int(num_iters), C.ptr(self.alpha),
                                                 C.ptr(self.beta), C.ptr(labels), C.ptr(ws), ws_bytes, C.stream_ptr(dev)),
                    "ssvb_sela_self_label")
        return labels
