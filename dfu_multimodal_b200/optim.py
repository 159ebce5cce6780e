"""Fused multi-tensor AdamW (a8): one kernel launch updates every parameter, its two moments and the bf16 weight
shadow the GEMMs read, with the step count held on the device so the update is CUDA-graph capturable.
Semantics follow torch.optim.AdamW as the reference configures it -- one param group over model.parameters(),
lr 1e-4, weight_decay 1e-4, betas (0.9, 0.999), eps 1e-8 (notebooks/train_multimodal_fusion.py:347)."""
import ctypes

import torch

from . import _lib, ops
from ._lib import AdamWTensor, check
from .functional import mark_shadow_fresh

CHUNK = 65536


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._tables = {}      # group index -> cached device table
        self._grad_source = None
        self._steps = {}

    def set_grad_source(self, source):
        """source(p) -> (tensor_with_grad_in_storage_order, is_bf16, scale): lets the data-parallel bucketer hand
        the all-reduced bf16 buckets straight to the update kernel."""
        self._grad_source = source
        self._tables.clear()

    def _state_for(self, p):
        st = self.state[p]
        if len(st) == 0:
            key = (p.device.type, p.device.index)
            if key not in self._steps:
                self._steps[key] = torch.zeros((), dtype=torch.float32, device=p.device)
            st["step"] = self._steps[key]      # one device-resident counter shared by all parameters
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _build(self, gi, group, params, grads, g_bf16):
        dev = params[0].device
        n = len(params)
        arr = (AdamWTensor * n)()
        chunk_tensor, chunk_offset = [], []
        shadowed = []
        for i, (p, g) in enumerate(zip(params, grads)):
            st = self._state_for(p)
            sh = getattr(p, "_dfu_shadow", None)
            shadow_ptr = None
            if sh is not None and sh[1].numel() == p.numel() and sh[1].device == p.device:
                shadow_ptr = sh[1].data_ptr()
                shadowed.append(p)
            arr[i].p, arr[i].g = p.data_ptr(), g.data_ptr()
            arr[i].m, arr[i].v = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
            arr[i].shadow, arr[i].numel = shadow_ptr, p.numel()
            for c in range((p.numel() + CHUNK - 1) // CHUNK):
                chunk_tensor.append(i)
                chunk_offset.append(c)
        # pinned staging + async copies: legal inside CUDA-graph capture (they become memcpy nodes whose pinned
        # sources are kept alive in the cache entry)
        host = [torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).pin_memory(),
                torch.tensor(chunk_tensor, dtype=torch.int32).pin_memory(),
                torch.tensor(chunk_offset, dtype=torch.int32).pin_memory()]
        table, ct, co = [h.to(dev, non_blocking=True) for h in host]
        sig = tuple((p.data_ptr(), g.data_ptr(), getattr(p, "_dfu_shadow", (None, None))[1] is not None)
                    for p, g in zip(params, grads))
        # one shared device step counter per group (all params step together)
        step = self.state[params[0]]["step"]
        return dict(table=table, ct=ct, co=co, host=host, n=n, n_chunks=len(chunk_tensor), sig=sig, shadowed=shadowed, step=step,
                    g_bf16=g_bf16)

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        L = _lib.load()
        st = torch.cuda.current_stream().cuda_stream
        for gi, group in enumerate(self.param_groups):
            params, grads = [], []
            g_bf16, scale = False, 1.0
            for p in group["params"]:
                if self._grad_source is not None:
                    src = self._grad_source(p)
                    if src is None:
                        continue
                    g, g_bf16, scale = src
                else:
                    g = p.grad
                    if g is None:
                        continue
                    # same element order in memory?  strides of size-1 dims are arbitrary (a 1x1 convolution's weight
                    # gradient arrives as (C,1,C,C) against the parameter's (C,1,1,1)) and must not force a copy
                    same_order = all(gs == ps for gs, ps, n in zip(g.stride(), p.stride(), p.shape) if n > 1)
                    if g.dtype != torch.float32 or g.shape != p.shape or not same_order:
                        g2 = torch.empty_like(p, memory_format=torch.preserve_format)
                        g2.copy_(g)
                        g = g2
                if not p.is_cuda:
                    raise RuntimeError("FusedAdamW has no CPU path; parameter lives on %s" % p.device)
                params.append(p)
                grads.append(g)
            if not params:
                continue
            cached = self._tables.get(gi)
            sig = tuple((p.data_ptr(), g.data_ptr(), getattr(p, "_dfu_shadow", (None, None))[1] is not None)
                        for p, g in zip(params, grads))
            if cached is None or cached["sig"] != sig:
                cached = self._build(gi, group, params, grads, g_bf16)
                self._tables[gi] = cached
            b1, b2 = group["betas"]
            check(L.dfu_adamw_tick(cached["step"].data_ptr(), st), "adamw_tick")
            check(L.dfu_adamw(cached["table"].data_ptr(), cached["n"], cached["ct"].data_ptr(), cached["co"].data_ptr(),
                              cached["n_chunks"], CHUNK, cached["step"].data_ptr(), group["lr"], b1, b2, group["eps"],
                              group["weight_decay"], scale, 1 if cached["g_bf16"] else 0, st), "adamw")
            ops._count(2)
            fresh = set(id(p) for p in cached["shadowed"])
            for p in params:
                if id(p) in fresh:
                    mark_shadow_fresh(p)
                elif getattr(p, "_dfu_shadow", None) is not None:
                    p._dfu_shadow = (None, p._dfu_shadow[1])     # padded shadow: recast on next use
        return loss
