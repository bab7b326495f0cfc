"""Drop-in loss modules: same names / ctor kwargs / forward signatures as the reference
`utils/losses.py` (SimclrLoss :8-46, MocoLoss :49-72, BarlowLoss :120-142, SimSiamLoss :145-151,
RelicLoss :154-201, SwavLoss :204-235) plus an nn.MSELoss()-compatible module for BYOL
(models/byol.py:89).  Each forward returns a 0-dim tensor with autograd history; forward and
backward both run in the CUDA kernels of libssv_b200.so through ctypes (no torch ops on the path
except buffer allocation).
"""
from __future__ import annotations

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import _cabi as C


def _ld(t):
    return t.stride(0)


# --------------------------------------------------------------------------------------------- NT-Xent
class _NtxentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, zi, zj, normalize, temperature):
        C.require_cuda(zi, zj)
        if zi.shape != zj.shape or zi.dim() != 2:
            raise ValueError("SimclrLoss expects two [N, d] tensors of the same shape")
        xi, xj = C.as_f32_rows(zi), C.as_f32_rows(zj)
        n, d = xi.shape
        L = C.lib()
        dev = xi.device
        with C.on_device(dev):
            saved = C.byte_buffer(C.cached_size("ssvb_ntxent_saved_bytes", n, d), dev)
            ws_bytes = C.cached_size("ssvb_ntxent_workspace_bytes", n, d)
            ws = C.workspace("ntxent", ws_bytes, dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            C.check(L.ssvb_ntxent_fwd(C.ptr(xi), C.ptr(xj), n, d, _ld(xi), _ld(xj), int(bool(normalize)),
                                      float(temperature), C.ptr(loss), C.ptr(saved), C.ptr(ws), ws_bytes,
                                      C.stream_ptr(dev)), "ssvb_ntxent_fwd")
        ctx.save_for_backward(xi, xj, saved)
        ctx.cfg = (int(bool(normalize)), float(temperature), zi.dtype, zj.dtype)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        xi, xj, saved = ctx.saved_tensors
        normalize, temperature, dti, dtj = ctx.cfg
        n, d = xi.shape
        L = C.lib()
        dev = xi.device
        with C.on_device(dev):
            go = C.f32_scalar(grad_out)
            dzi, dzj = torch.empty_like(xi), torch.empty_like(xj)
            ws_bytes = C.cached_size("ssvb_ntxent_workspace_bytes", n, d)
            ws = C.workspace("ntxent", ws_bytes, dev)
            C.check(L.ssvb_ntxent_bwd(C.ptr(xi), C.ptr(xj), n, d, _ld(xi), _ld(xj), normalize, temperature,
                                      C.ptr(go), C.ptr(saved), C.ptr(dzi), C.ptr(dzj), _ld(dzi), _ld(dzj),
                                      C.ptr(ws), ws_bytes, C.stream_ptr(dev)), "ssvb_ntxent_bwd")
        return dzi.to(dti), dzj.to(dtj), None, None


class SimclrLoss(nn.Module):
    """Reference: utils/losses.py:8-46 (ctor defaults normalize=False, temperature=1.0)."""

    def __init__(self, normalize=False, temperature=1.0):
        super().__init__()
        self.normalize = normalize
        self.temperature = temperature

    def forward(self, zi, zj):
        return _NtxentFn.apply(zi, zj, self.normalize, self.temperature)


# --------------------------------------------------------------------------------------------- MoCo
def _bank_shadow(memory_vectors):
    """bf16 shadow of a device-resident MemoryBank (see banks.py) if `memory_vectors` is exactly its
    live fp32 storage; otherwise None (the fp32 queue is converted inside the C call)."""
    from .banks import lookup_shadow
    return lookup_shadow(memory_vectors)


class _MocoFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, query, keys, memory_vectors, normalize, temperature):
        C.require_cuda(query, keys, memory_vectors)
        if query.shape != keys.shape or query.dim() != 2 or memory_vectors.dim() != 2 \
                or memory_vectors.shape[1] != query.shape[1]:
            raise ValueError("MocoLoss expects query/keys [N, d] and memory_vectors [K, d]")
        q, k = C.as_f32_rows(query), C.as_f32_rows(keys)
        mem = C.as_f32_rows(memory_vectors.detach())
        shadow = _bank_shadow(memory_vectors) if mem.data_ptr() == memory_vectors.data_ptr() else None
        n, d = q.shape
        kq = mem.shape[0]
        L = C.lib()
        dev = q.device
        with C.on_device(dev):
            saved = C.byte_buffer(C.cached_size("ssvb_moco_saved_bytes", n, kq, d), dev)
            ws_bytes = C.cached_size("ssvb_moco_workspace_bytes", n, kq, d)
            ws = C.workspace("moco", ws_bytes, dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            # a MemoryBank's rows are unit-norm or zero by construction (normalised on enqueue): with its shadow in hand the
            # fused single-pass kernel applies (queue read once, backward without queue access)
            unit = int(shadow is not None)
            C.check(L.ssvb_moco_fwd(C.ptr(q), C.ptr(k), C.ptr(mem), C.ptr(shadow), unit, n, kq, d, _ld(q), _ld(k), _ld(mem),
                                    int(bool(normalize)), float(temperature), C.ptr(loss), C.ptr(saved), C.ptr(ws),
                                    ws_bytes, C.stream_ptr(dev)), "ssvb_moco_fwd")
        # fused single-pass form (csrc/moco.cu `moco_fused`): everything the backward needs was produced by the forward
        # and lives in `saved`, so the queue is NOT a saved tensor - the bank may be enqueued before backward (any loop
        # order works, like the reference whose loss sees a per-step copy of the queue, models/moco.py:117).  Otherwise
        # the backward re-reads the queue and `mem` is saved, so autograd's version check guards it.
        fused = bool(unit and normalize and 2.0 * 1.4426950408889634 / float(temperature) <= 120.0)
        if fused:
            ctx.save_for_backward(q, k, saved)
            ctx.mem = mem
        else:
            ctx.save_for_backward(q, k, mem, saved)
        ctx.shadow = shadow
        ctx.cfg = (int(bool(normalize)), float(temperature), query.dtype, keys.dtype)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        if len(ctx.saved_tensors) == 3:
            (q, k, saved), mem = ctx.saved_tensors, ctx.mem
        else:
            q, k, mem, saved = ctx.saved_tensors
        normalize, temperature, dtq, dtk = ctx.cfg
        n, d = q.shape
        kq = mem.shape[0]
        L = C.lib()
        dev = q.device
        with C.on_device(dev):
            go = C.f32_scalar(grad_out)
            dq, dk = torch.empty_like(q), torch.empty_like(k)
            ws_bytes = C.cached_size("ssvb_moco_workspace_bytes", n, kq, d)
            ws = C.workspace("moco", ws_bytes, dev)
            C.check(L.ssvb_moco_bwd(C.ptr(q), C.ptr(k), C.ptr(mem), C.ptr(ctx.shadow), int(ctx.shadow is not None), n, kq, d,
                                    _ld(q), _ld(k),
                                    _ld(mem), normalize, temperature, C.ptr(go), C.ptr(saved), C.ptr(dq), C.ptr(dk),
                                    _ld(dq), _ld(dk), C.ptr(ws), ws_bytes, C.stream_ptr(dev)), "ssvb_moco_bwd")
        return dq.to(dtq), dk.to(dtk), None, None, None


class MocoLoss(nn.Module):
    """Reference: utils/losses.py:49-72 (ctor defaults normalize=True, temperature=1.0)."""

    def __init__(self, normalize=True, temperature=1.0):
        super().__init__()
        self.normalize = normalize
        self.temperature = temperature

    def forward(self, query, keys, memory_vectors):
        return _MocoFn.apply(query, keys, memory_vectors, self.normalize, self.temperature)


# --------------------------------------------------------------------------------------------- row-dot losses
class _RowdotFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, o, t, kind):
        C.require_cuda(o, t)
        if o.shape != t.shape:
            raise ValueError("row-dot losses expect two tensors of the same shape")
        if kind == 1 and o.dim() != 2:
            # -(o * t).sum(1).mean() (utils/losses.py:150-151) sums over dim 1: for > 2-D inputs that is not the last
            # axis and the divisor differs from a flattened [N, d] view; the reference only ever passes [N, d]
            raise ValueError("SimSiamLoss expects [N, d] inputs (the reference's shape); reshape before the call")
        oo = C.as_f32_rows(o.reshape(-1, o.shape[-1]) if o.dim() != 2 else o)
        tt = C.as_f32_rows(t.reshape(-1, t.shape[-1]) if t.dim() != 2 else t)
        n, d = oo.shape
        L = C.lib()
        dev = oo.device
        with C.on_device(dev):
            ws_bytes = C.cached_size("ssvb_rowdot_workspace_bytes", n, d)
            ws = C.workspace("rowdot", ws_bytes, dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            C.check(L.ssvb_rowdot_fwd(kind, C.ptr(oo), C.ptr(tt), n, d, _ld(oo), _ld(tt), C.ptr(loss), C.ptr(ws),
                                      ws_bytes, C.stream_ptr(dev)), "ssvb_rowdot_fwd")
        ctx.save_for_backward(oo, tt)
        ctx.cfg = (kind, o.shape, o.dtype, t.dtype)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        oo, tt = ctx.saved_tensors
        kind, shape, dto, dtt = ctx.cfg
        need_o, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_o or need_t):
            return None, None, None
        n, d = oo.shape
        L = C.lib()
        dev = oo.device
        with C.on_device(dev):
            go = C.f32_scalar(grad_out)
            d_o = torch.empty_like(oo) if need_o else None
            d_t = torch.empty_like(tt) if need_t else None
            C.check(L.ssvb_rowdot_bwd(kind, C.ptr(oo), C.ptr(tt), n, d, _ld(oo), _ld(tt), C.ptr(go), C.ptr(d_o),
                                      C.ptr(d_t), _ld(d_o) if need_o else 0, _ld(d_t) if need_t else 0,
                                      C.stream_ptr(dev)), "ssvb_rowdot_bwd")
        return (d_o.reshape(shape).to(dto) if need_o else None,
                d_t.reshape(shape).to(dtt) if need_t else None, None)


class SimSiamLoss(nn.Module):
    """Reference: utils/losses.py:145-151 : -(online * target).sum(1).mean()."""

    def __init__(self):
        super().__init__()

    def forward(self, online_output, target_output):
        return _RowdotFn.apply(online_output, target_output, 1)


class MSELoss(nn.Module):
    """nn.MSELoss()-compatible (mean reduction) for BYOL: models/byol.py:89, used :129-130."""

    def __init__(self):
        super().__init__()

    def forward(self, input, target):  # noqa: A002  (same parameter names as torch.nn.MSELoss)
        return _RowdotFn.apply(input, target, 0)


# --------------------------------------------------------------------------------------------- ReLIC
class _RelicFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, zi, zj, zo, normalize, temperature, alpha):
        C.require_cuda(zi, zj, zo)
        if not (zi.shape == zj.shape == zo.shape) or zi.dim() != 2:
            raise ValueError("RelicLoss expects three [N, d] tensors of the same shape")
        xi, xj, xo = C.as_f32_rows(zi), C.as_f32_rows(zj), C.as_f32_rows(zo)
        n, d = xi.shape
        L = C.lib()
        dev = xi.device
        norm = int(bool(normalize))
        with C.on_device(dev):
            st = C.stream_ptr(dev)
            saved_c = C.byte_buffer(C.cached_size("ssvb_ntxent_saved_bytes", n, d), dev)
            ws_bytes = C.cached_size("ssvb_ntxent_workspace_bytes", n, d)
            ws = C.workspace("ntxent", ws_bytes, dev)
            out = torch.empty((), dtype=torch.float32, device=dev)
            saved_k = C.byte_buffer(C.cached_size("ssvb_relic_kl_saved_bytes", n), dev)
            C.check(L.ssvb_relic_fwd(C.ptr(xi), C.ptr(xj), C.ptr(xo), n, d, _ld(xi), _ld(xj), _ld(xo), norm,
                                     float(temperature), float(alpha), C.ptr(out), C.ptr(saved_c), C.ptr(saved_k),
                                     C.ptr(ws), ws_bytes, st), "ssvb_relic_fwd")
        ctx.save_for_backward(xi, xj, xo, saved_c, saved_k)
        ctx.cfg = (norm, float(temperature), float(alpha), zi.dtype, zj.dtype, zo.dtype)
        return out  # contrastive + alpha * KL  (utils/losses.py:201), combined by the last kernel

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        xi, xj, xo, saved_c, saved_k = ctx.saved_tensors
        norm, temperature, alpha, dti, dtj, dto = ctx.cfg
        n, d = xi.shape
        L = C.lib()
        dev = xi.device
        with C.on_device(dev):
            st = C.stream_ptr(dev)
            go = C.f32_scalar(grad_out)
            dzi, dzj, dzo = torch.empty_like(xi), torch.empty_like(xj), torch.empty_like(xo)
            ws_bytes = C.cached_size("ssvb_ntxent_workspace_bytes", n, d)
            ws = C.workspace("ntxent", ws_bytes, dev)
            C.check(L.ssvb_relic_bwd(C.ptr(xi), C.ptr(xj), C.ptr(xo), n, d, _ld(xi), _ld(xj), _ld(xo), norm, temperature,
                                     alpha, C.ptr(go), C.ptr(saved_c), C.ptr(saved_k), C.ptr(dzi), C.ptr(dzj), C.ptr(dzo),
                                     _ld(dzi), _ld(dzj), _ld(dzo), C.ptr(ws), ws_bytes, st), "ssvb_relic_bwd")
        return dzi.to(dti), dzj.to(dtj), dzo.to(dto), None, None, None


class RelicLoss(nn.Module):
    """Reference: utils/losses.py:154-201 (ctor defaults normalize=True, temperature=1.0, alpha=0.5)."""

    def __init__(self, normalize=True, temperature=1.0, alpha=0.5):
        super().__init__()
        self.normalize = normalize
        self.temperature = temperature
        self.alpha = alpha

    def forward(self, zi, zj, z_orig):
        return _RelicFn.apply(zi, zj, z_orig, self.normalize, self.temperature, self.alpha)


# --------------------------------------------------------------------------------------------- Barlow Twins
class _BarlowFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, zi, zj, normalize, lmbda):
        C.require_cuda(zi, zj)
        if zi.shape != zj.shape or zi.dim() != 2:
            raise ValueError("BarlowLoss expects two [N, D] tensors of the same shape")
        xi, xj = C.as_f32_rows(zi), C.as_f32_rows(zj)
        n, d = xi.shape
        L = C.lib()
        dev = xi.device
        with C.on_device(dev):
            saved = C.byte_buffer(C.cached_size("ssvb_barlow_saved_bytes", n, d), dev)
            ws_bytes = C.cached_size("ssvb_barlow_workspace_bytes", n, d)
            ws = C.workspace("barlow", ws_bytes, dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            C.check(L.ssvb_barlow_fwd(C.ptr(xi), C.ptr(xj), n, d, _ld(xi), _ld(xj), int(bool(normalize)), float(lmbda),
                                      C.ptr(loss), C.ptr(saved), C.ptr(ws), ws_bytes, C.stream_ptr(dev)),
                    "ssvb_barlow_fwd")
        ctx.save_for_backward(xi, xj, saved)
        ctx.cfg = (int(bool(normalize)), float(lmbda), zi.dtype, zj.dtype)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        xi, xj, saved = ctx.saved_tensors
        normalize, lmbda, dti, dtj = ctx.cfg
        n, d = xi.shape
        L = C.lib()
        dev = xi.device
        with C.on_device(dev):
            go = C.f32_scalar(grad_out)
            dzi, dzj = torch.empty_like(xi), torch.empty_like(xj)
            ws_bytes = C.cached_size("ssvb_barlow_workspace_bytes", n, d)
            ws = C.workspace("barlow", ws_bytes, dev)
            C.check(L.ssvb_barlow_bwd(C.ptr(xi), C.ptr(xj), n, d, _ld(xi), _ld(xj), normalize, lmbda, C.ptr(go),
                                      C.ptr(saved), C.ptr(dzi), C.ptr(dzj), _ld(dzi), _ld(dzj), C.ptr(ws), ws_bytes,
                                      C.stream_ptr(dev)), "ssvb_barlow_bwd")
        return dzi.to(dti), dzj.to(dtj), None, None


class BarlowLoss(nn.Module):
    """Reference: utils/losses.py:120-142 (ctor defaults normalize=True, off_diagonal_weight=0.005)."""

    def __init__(self, normalize=True, off_diagonal_weight=0.005):
        super().__init__()
        self.normalize = normalize
        self.lmbda = off_diagonal_weight

    def forward(self, z_i, z_j):
        return _BarlowFn.apply(z_i, z_j, self.normalize, self.lmbda)


# --------------------------------------------------------------------------------------------- SwAV
class _SwavFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, prototypes, bank, temperature, eps, n_iters):
        C.require_cuda(z1, z2, prototypes, bank)
        if z1.shape != z2.shape or z1.dim() != 2 or prototypes.dim() != 2 or prototypes.shape[1] != z1.shape[1]:
            raise ValueError("SwavLoss expects z_1/z_2 [B, d] and prototypes [K, d]")
        x1, x2, pc = C.as_f32_rows(z1), C.as_f32_rows(z2), C.as_f32_rows(prototypes)
        bk = C.as_f32_rows(bank.detach()) if bank is not None and bank.shape[0] > 0 else None
        nb, d = x1.shape
        nbank = bk.shape[0] if bk is not None else 0
        k = pc.shape[0]
        L = C.lib()
        dev = x1.device
        with C.on_device(dev):
            saved = C.byte_buffer(C.cached_size("ssvb_swav_saved_bytes", nb, nbank, k, d), dev)
            ws_bytes = C.cached_size("ssvb_swav_workspace_bytes", nb, nbank, k, d)
            ws = C.workspace("swav", ws_bytes, dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            C.check(L.ssvb_swav_fwd(C.ptr(x1), C.ptr(x2), C.ptr(bk), C.ptr(pc), nb, nbank, k, d, _ld(x1), _ld(x2),
                                    _ld(bk) if bk is not None else 0, _ld(pc), float(temperature), float(eps),
                                    int(n_iters), C.ptr(loss), C.ptr(saved), C.ptr(ws), ws_bytes, C.stream_ptr(dev)),
                    "ssvb_swav_fwd")
        ctx.save_for_backward(x1, x2, pc, saved, *([bk] if bk is not None else []))
        ctx.cfg = (float(temperature), z1.dtype, z2.dtype, prototypes.dtype)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        x1, x2, pc, saved, *rest = ctx.saved_tensors
        bk = rest[0] if rest else None
        temperature, dt1, dt2, dtp = ctx.cfg
        nb, d = x1.shape
        nbank = bk.shape[0] if bk is not None else 0
        k = pc.shape[0]
        L = C.lib()
        dev = x1.device
        with C.on_device(dev):
            go = C.f32_scalar(grad_out)
            dz1, dz2, dpc = torch.empty_like(x1), torch.empty_like(x2), torch.empty_like(pc)
            ws_bytes = C.cached_size("ssvb_swav_workspace_bytes", nb, nbank, k, d)
            ws = C.workspace("swav", ws_bytes, dev)
            C.check(L.ssvb_swav_bwd(C.ptr(x1), C.ptr(x2), C.ptr(bk), C.ptr(pc), nb, nbank, k, d, _ld(x1), _ld(x2),
                                    _ld(bk) if bk is not None else 0, _ld(pc), temperature, C.ptr(go), C.ptr(saved),
                                    C.ptr(dz1), C.ptr(dz2), C.ptr(dpc), _ld(dz1), _ld(dz2), _ld(dpc), C.ptr(ws),
                                    ws_bytes, C.stream_ptr(dev)), "ssvb_swav_bwd")
        return dz1.to(dt1), dz2.to(dt2), dpc.to(dtp), None, None, None, None


class SwavLoss(nn.Module):
    """Reference: utils/losses.py:204-235 (ctor defaults temperature=0.1, sinkhorn_eps=0.05, sinkhorn_iters=3).
    `device` is kept as an attribute because the reference exposes one (utils/losses.py:208); it is unused."""

    def __init__(self, temperature=0.1, sinkhorn_eps=0.05, sinkhorn_iters=3):
        super().__init__()
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.temperature = temperature
        self.n_iters = sinkhorn_iters
        self.eps = sinkhorn_eps

    @torch.no_grad()
    def compute_codes_sinkhorn(self, scores):
        C.require_cuda(scores)
        if scores.dim() != 2:
            raise ValueError("compute_codes_sinkhorn expects [B, K] scores")
        s = scores.detach()
        if s.dtype != torch.float32 or s.stride(1) != 1:
            s = s.float().contiguous()
        b, k = s.shape
        L = C.lib()
        dev = s.device
        with C.on_device(dev):
            codes = torch.empty(b, k, dtype=torch.float32, device=dev)
            ws_bytes = C.cached_size("ssvb_sinkhorn_workspace_bytes", b, k)
            ws = C.workspace("sinkhorn", ws_bytes, dev)
            C.check(L.ssvb_sinkhorn(C.ptr(s), b, k, s.stride(0), float(self.eps), int(self.n_iters), C.ptr(codes),
                                    codes.stride(0), C.ptr(ws), ws_bytes, C.stream_ptr(dev)), "ssvb_sinkhorn")
        return codes

    def forward(self, z_1, z_2, prototypes, bank_features=None):
        return _SwavFn.apply(z_1, z_2, prototypes, bank_features, self.temperature, self.eps, self.n_iters)


# --------------------------------------------------------------------------------------------- DINO
class _DinoFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, teacher, student, temp_s, temp_t, center):
        C.require_cuda(teacher, student, center)
        if teacher.dim() != 3 or student.dim() != 3 or teacher.shape[1] != 2 or teacher.shape[0] != student.shape[0] \
                or teacher.shape[2] != student.shape[2] or center.numel() != student.shape[2]:
            raise ValueError("DinoLoss expects teacher [bs, 2, K], student [bs, 2+V, K] and center [K]")
        t = teacher.detach().float().contiguous()
        s = student.detach().float().contiguous()
        c = center.detach().float().contiguous().view(-1)
        bs, nv, k = s.shape
        L = C.lib()
        dev = s.device
        with C.on_device(dev):
            ws_bytes = C.cached_size("ssvb_dino_workspace_bytes", bs, nv, k)
            ws = C.workspace("dino", ws_bytes, dev)
            loss = torch.empty((), dtype=torch.float32, device=dev)
            C.check(L.ssvb_dino_fwd(C.ptr(t), C.ptr(s), C.ptr(c), bs, nv, k, float(temp_s), float(temp_t), C.ptr(loss),
                                    C.ptr(ws), ws_bytes, C.stream_ptr(dev)), "ssvb_dino_fwd")
        ctx.save_for_backward(t, s, c)
        ctx.cfg = (float(temp_s), float(temp_t), student.dtype)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        t, s, c = ctx.saved_tensors
        temp_s, temp_t, dts = ctx.cfg
        bs, nv, k = s.shape
        dev = s.device
        with C.on_device(dev):
            go = C.f32_scalar(grad_out)
            ds = torch.empty_like(s)
            ws_bytes = C.cached_size("ssvb_dino_workspace_bytes", bs, nv, k)
            ws = C.workspace("dino", ws_bytes, dev)
            C.check(C.lib().ssvb_dino_bwd(C.ptr(t), C.ptr(s), C.ptr(c), bs, nv, k, temp_s, temp_t, C.ptr(go), C.ptr(ds),
                                          C.ptr(ws), ws_bytes, C.stream_ptr(dev)), "ssvb_dino_bwd")
        return None, ds.to(dts), None, None, None


class DinoLoss(nn.Module):
    """Reference: utils/losses.py:75-89 (call site models/dino.py:161-162).  Gradient flows to the student only (the
    reference evaluates the teacher under no_grad, models/dino.py:151-152)."""

    def __init__(self):
        super().__init__()

    def forward(self, teacher_fvecs, student_fvecs, temp_s, temp_t, center):
        return _DinoFn.apply(teacher_fvecs, student_fvecs, temp_s, temp_t, center)


@torch.no_grad()
def update_teacher_center(center, teacher_fvecs, momentum):
    """models/dino.py:136-141: returns the new centre (`center` may be None on the first call, as in the reference)."""
    C.require_cuda(teacher_fvecs, center)
    t = C.as_f32_rows(teacher_fvecs.detach().reshape(-1, teacher_fvecs.shape[-1]))
    rows, k = t.shape
    first = center is None
    out = torch.empty(k, dtype=torch.float32, device=t.device) if first else center.detach().float().contiguous().clone()
    with C.on_device(t.device):
        C.check(C.lib().ssvb_dino_center_update(C.ptr(t), rows, k, t.stride(0), float(momentum), float(1.0 - momentum),
                                                int(first), C.ptr(out), C.stream_ptr(t.device)),
                "ssvb_dino_center_update")
    return out


# --------------------------------------------------------------------------------------------- PIRL
class _PirlFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, patch, mem_pos, mem_neg, normalize, temperature, loss_weight):
        C.require_cuda(img, patch, mem_pos, mem_neg)
        if img.shape != patch.shape or img.shape != mem_pos.shape or img.dim() != 2 or mem_neg.dim() != 2 \
                or mem_neg.shape[1] != img.shape[1]:
            raise ValueError("PirlLoss expects img/patch/memory_pos [N, d] and memory_neg [K, d]")
        xi, xp = C.as_f32_rows(img), C.as_f32_rows(patch)
        mp, mn = C.as_f32_rows(mem_pos.detach()), C.as_f32_rows(mem_neg.detach())
        n, d = xi.shape
        k = mn.shape[0]
        L = C.lib()
        dev = xi.device
        norm = int(bool(normalize))
        with C.on_device(dev):
            saved = C.byte_buffer(C.cached_size("ssvb_pirl_saved_bytes", n, d), dev)
            ws_bytes = C.cached_size("ssvb_pirl_workspace_bytes", n, k, d)
            ws = C.workspace("moco", ws_bytes, dev)
            out = torch.empty((), dtype=torch.float32, device=dev)
            C.check(L.ssvb_pirl_fwd(C.ptr(xi), C.ptr(xp), C.ptr(mp), C.ptr(mn), n, k, d, _ld(xi), _ld(xp), _ld(mp), _ld(mn),
                                    norm, float(temperature), float(loss_weight), C.ptr(out), C.ptr(saved), C.ptr(ws),
                                    ws_bytes, C.stream_ptr(dev)), "ssvb_pirl_fwd")
        ctx.save_for_backward(xi, xp, mp, saved)
        ctx.cfg = (norm, float(temperature), float(loss_weight), img.dtype, patch.dtype)
        return out  # loss_weight * loss_1 + (1 - loss_weight) * loss_2  (utils/losses.py:117), summed in the kernel

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        xi, xp, mp, saved = ctx.saved_tensors
        norm, temperature, w, dti, dtp = ctx.cfg
        n, d = xi.shape
        dev = xi.device
        with C.on_device(dev):
            go = C.f32_scalar(grad_out)
            di, dp = torch.empty_like(xi), torch.empty_like(xp)
            C.check(C.lib().ssvb_pirl_bwd(C.ptr(xi), C.ptr(xp), C.ptr(mp), n, d, _ld(xi), _ld(xp), _ld(mp), norm,
                                          temperature, w, C.ptr(go), C.ptr(saved), C.ptr(di), C.ptr(dp), _ld(di), _ld(dp),
                                          C.stream_ptr(dev)), "ssvb_pirl_bwd")
        return di.to(dti), dp.to(dtp), None, None, None, None, None


class PirlLoss(nn.Module):
    """Reference: utils/losses.py:92-117 (ctor defaults normalize=True, temperature=1.0, loss_weight=0.5)."""

    def __init__(self, normalize=True, temperature=1.0, loss_weight=0.5):
        super().__init__()
        self.loss_weight = loss_weight
        self.normalize = normalize
        self.temp = temperature

    def forward(self, img_features, patch_features, memory_pos_features, memory_neg_features):
        return _PirlFn.apply(img_features, patch_features, memory_pos_features, memory_neg_features, self.normalize,
                             self.temp, self.loss_weight)
