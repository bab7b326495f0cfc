"""Device-resident ring buffers and prototypes with the reference's attribute / method names.

MemoryBank   — models/moco.py:23-39   (`bank`, `size`, `ptr`, `add_batch`, `get_vectors`)
FeatureBank  — models/swav.py:57-79   (`vectors`, `bank_size`, `ptr`, `add_vectors`, `return_vectors`)
Prototypes   — models/swav.py:44-54   (`embedding`, `proto_size`, `forward(device)`)

The reference keeps the banks on the CPU, writes them row by row in a Python loop and copies the whole
bank host->device every step; here they live in HBM and one enqueue kernel writes all rows.  Pointer
bookkeeping (`ptr`) stays a host-side Python int and is bit-exact with the reference loop.
"""
from __future__ import annotations

import ctypes
import weakref

import torch
import torch.nn as nn
from torch.autograd.function import once_differentiable

from . import _cabi as C

_SHADOWS = {}  # data_ptr of a bank's fp32 storage -> weakref to its owner


def lookup_shadow(t):
    ref = _SHADOWS.get(t.data_ptr())
    owner = ref() if ref is not None else None
    if owner is None or owner._shadow is None:
        return None
    live = owner._storage()
    if live is not t and (live.data_ptr() != t.data_ptr() or live.shape != t.shape):
        return None
    if live._version != owner._shadow_version:  # someone wrote the bank with torch ops: shadow is stale
        return None
    return owner._shadow


def _default_device(device):
    if device is not None:
        return torch.device(device)
    if not torch.cuda.is_available():
        raise RuntimeError("ssv_b200 banks live in B200 HBM: no CUDA device available and there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class _Ring:
    def _init_ring(self, rows, dim, device, normalize, with_shadow):
        self._device = _default_device(device)
        self._normalize = normalize
        self._data = torch.zeros(rows, dim, dtype=torch.float32, device=self._device)
        self._shadow = None
        self._shadow_version = -1
        if with_shadow and dim % 4 == 0 and dim <= 256:
            dpad = C.lib().ssvb_ntxent_dpad(dim)
            self._shadow = torch.zeros(rows, dpad, dtype=torch.bfloat16, device=self._device)
            self._shadow_version = self._data._version
            _SHADOWS[self._data.data_ptr()] = weakref.ref(self)
        self.ptr = 0

    def _storage(self):
        return self._data

    def _enqueue(self, batch):
        if batch.dim() != 2 or batch.shape[1] != self._data.shape[1]:
            raise ValueError(f"expected a [N, {self._data.shape[1]}] batch")
        b = C.as_f32_rows(batch.detach().to(self._device, non_blocking=True))
        rows, dim = self._data.shape
        new_ptr = ctypes.c_int64(-1)
        shadow_ok = self._shadow is not None and self._shadow_version == self._data._version
        with torch.cuda.device(self._device):
            C.check(C.lib().ssvb_ring_enqueue(C.ptr(self._data), C.ptr(self._shadow) if shadow_ok else None, rows, dim,
                                              self._data.stride(0), C.ptr(b), b.shape[0], b.stride(0), int(self.ptr),
                                              int(self._normalize), ctypes.cast(ctypes.pointer(new_ptr), ctypes.c_void_p),
                                              C.stream_ptr(self._device)), "ssvb_ring_enqueue")
        self.ptr = int(new_ptr.value)
        # the kernel wrote the storage behind torch's back: bump the version counter so autograd's saved-tensor check
        # fires if a pending backward still needs the OLD contents (MocoLoss re-reads the queue in backward; the
        # reference is immune because `get_vectors().to(device)` hands it a copy), then re-sync the shadow's version
        torch.autograd.graph.increment_version(self._data)
        if shadow_ok:
            self._shadow_version = self._data._version


class MemoryBank(_Ring):
    """MoCo queue: rows are L2-normalised on enqueue; starts as zeros (normalize(0) == 0)."""

    def __init__(self, queue_size, feature_size, device=None):
        self._init_ring(queue_size, feature_size, device, normalize=True, with_shadow=True)
        self.size = queue_size

    @property
    def bank(self):
        return self._data

    def add_batch(self, batch):
        self._enqueue(batch)

    def get_vectors(self):
        return self._data


class FeatureBank(_Ring):
    """SwAV feature bank: rows stored as given (no normalisation)."""

    def __init__(self, bank_size, feature_dim, device=None):
        self._init_ring(bank_size, feature_dim, device, normalize=False, with_shadow=False)
        self.bank_size = bank_size

    @property
    def vectors(self):
        return self._data

    def add_vectors(self, fvecs):
        self._enqueue(fvecs)

    def return_vectors(self, device):
        # a SNAPSHOT, like the reference's host->device copy (models/swav.py:77-79): its own training loop enqueues the
        # new features BEFORE loss.backward() (swav.py:140-144), so the loss must not alias the live ring
        return self._data.to(device, copy=True)


class _L2NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        C.require_cuda(x)
        xx = C.as_f32_rows(x)
        n, d = xx.shape
        dev = xx.device
        with torch.cuda.device(dev):
            y = torch.empty_like(xx)
            inv = torch.empty(n, dtype=torch.float32, device=dev)
            C.check(C.lib().ssvb_l2norm_fwd(C.ptr(xx), n, d, xx.stride(0), C.ptr(y), y.stride(0), C.ptr(inv),
                                            C.stream_ptr(dev)), "ssvb_l2norm_fwd")
        ctx.save_for_backward(y, inv)
        ctx.dtype = x.dtype
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        y, inv = ctx.saved_tensors
        n, d = y.shape
        dev = y.device
        g = C.as_f32_rows(dy)
        with torch.cuda.device(dev):
            dx = torch.empty_like(y)
            C.check(C.lib().ssvb_l2norm_bwd(C.ptr(g), C.ptr(y), C.ptr(inv), n, d, g.stride(0), y.stride(0), C.ptr(dx),
                                            dx.stride(0), C.stream_ptr(dev)), "ssvb_l2norm_bwd")
        return dx.to(ctx.dtype)


def l2_normalize(x):
    """F.normalize(x, p=2, dim=-1) for a [N, d] CUDA tensor, forward and backward in ssv_b200 kernels."""
    return _L2NormFn.apply(x)


class Prototypes(nn.Module):
    """models/swav.py:44-54: embedding rows L2-normalised on every call (differentiable)."""

    def __init__(self, hidden_dim, prototype_size):
        super().__init__()
        self.proto_size = prototype_size
        self.embedding = nn.Embedding(prototype_size, hidden_dim)

    def forward(self, device):
        w = self.embedding.weight
        if w.device != torch.device(device):
            w = w.to(device)
        return l2_normalize(w)


class PirlMemoryBank:
    """Per-sample momentum bank of PIRL (reference models/pirl.py:22-46), device-resident: same attributes (`bank`,
    `num_negatives`, `data_size`, `size`, `m`, `ptr`) and methods.  initialize / update are one scatter kernel each
    (the reference round-trips through the CPU); the choice of negatives stays the reference's host-side randperm."""

    def __init__(self, data_size, feature_size, momentum=0.5, num_negatives=1000, device=None):
        self._device = _default_device(device)
        self.bank = torch.zeros(data_size, feature_size, dtype=torch.float32, device=self._device)
        self.num_negatives = num_negatives
        self.data_size = data_size
        self.size = data_size
        self.m = momentum
        self.ptr = 0

    def _idx(self, indices):
        return torch.as_tensor(indices).to(self._device, dtype=torch.int64).contiguous().view(-1)

    def _scatter(self, indices, vectors, mode):
        idx = self._idx(indices)
        v = C.as_f32_rows(vectors.detach().to(self._device))
        with torch.cuda.device(self._device):
            C.check(C.lib().ssvb_bank_scatter(C.ptr(self.bank), self.data_size, self.bank.shape[1], self.bank.stride(0),
                                              C.ptr(idx), idx.numel(), C.ptr(v), v.stride(0), float(self.m),
                                              float(1 - self.m), mode, C.stream_ptr(self._device)), "ssvb_bank_scatter")
        torch.autograd.graph.increment_version(self.bank)   # raw kernel write (get_positives / get_negatives return copies)

    def initialize_vectors(self, indices, vectors):
        self._scatter(indices, vectors, 0)

    def update_vectors(self, indices, new_vectors):
        self._scatter(indices, new_vectors, 1)

    def _gather(self, indices):
        idx = self._idx(indices)
        out = torch.empty(idx.numel(), self.bank.shape[1], dtype=torch.float32, device=self._device)
        with torch.cuda.device(self._device):
            C.check(C.lib().ssvb_bank_gather(C.ptr(self.bank), self.data_size, self.bank.shape[1], self.bank.stride(0),
                                             C.ptr(idx), idx.numel(), C.ptr(out), out.stride(0),
                                             C.stream_ptr(self._device)), "ssvb_bank_gather")
        return out

    def get_positives(self, indices):
        return self._gather(indices)

    def get_negatives(self, exclude_idx):
        # models/pirl.py:43-45: host-side random permutation minus the batch's own indices (bookkeeping, bit-exact
        # with the reference under the same torch RNG state)
        ex = set(int(i) for i in torch.as_tensor(exclude_idx).view(-1).tolist())
        indices = torch.tensor([i for i in torch.randperm(self.data_size).tolist() if i not in ex]).long()
        return self._gather(indices[:self.num_negatives])
