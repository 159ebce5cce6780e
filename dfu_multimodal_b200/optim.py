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
        # the extra keys are torch.optim.AdamW's own group entries with the values this class implements: a state_dict saved
        # here must make a stock AdamW behave as AdamW when it loads it (a group without `decoupled_weight_decay` is read as
        # plain Adam with L2 decay by torch >= 2.7)
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=False, maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None, decoupled_weight_decay=True)
        super().__init__(params, defaults)
        self._tables = {}      # (group index, pointer signature) -> cached device table
        self._grad_source = None
        self._bucketer = None
        self._steps = {}       # (device type, index) -> fp32 device scalar shared by every parameter on that device

    def set_grad_source(self, source, bucketer=None):
        """source(p) -> (tensor_with_grad_in_storage_order, is_bf16, scale): lets the data-parallel bucketer hand
        the all-reduced bf16 buckets straight to the update kernel.  With `bucketer` (ddp.GradBucketer) the update runs one
        launch per bucket, each waiting only for ITS bucket's all-reduce."""
        self._grad_source = source
        self._bucketer = bucketer
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
        # one shared device step counter per device (all params step together)
        step = self.state[params[0]]["step"]
        return dict(table=table, ct=ct, co=co, host=host, n=n, n_chunks=len(chunk_tensor), shadowed=shadowed, step=step,
                    g_bf16=g_bf16, captured=torch.cuda.is_current_stream_capturing())

    def _signature(self, params, grads):
        """Everything the device table bakes in: parameter, gradient, moment and shadow addresses."""
        sig = []
        for p, g in zip(params, grads):
            st = self.state.get(p, {})
            m, v = st.get("exp_avg"), st.get("exp_avg_sq")
            sh = getattr(p, "_dfu_shadow", None)
            sig.append((p.data_ptr(), g.data_ptr(), m.data_ptr() if m is not None else 0,
                        v.data_ptr() if v is not None else 0, sh[1].data_ptr() if sh is not None else 0))
        return tuple(sig)

    def _table_for(self, gi, group, params, grads, g_bf16):
        for p in params:
            self._state_for(p)
        key = (gi, self._signature(params, grads))
        cached = self._tables.get(key)
        if cached is None:
            cached = self._build(gi, group, params, grads, g_bf16)
            # tables built during CUDA-graph capture are baked into that graph (device table, pinned staging buffers) and
            # must outlive it; eager tables are keyed by allocator-dependent gradient addresses, keep only a few
            eager = [k for k, c in self._tables.items() if not c["captured"]]
            for k in eager[:max(0, len(eager) - 64)]:
                del self._tables[k]
            self._tables[key] = cached
        return cached

    # ---------------------------------------------------------------------------------------------------- checkpoints
    def state_dict(self):
        """Same schema as torch.optim.AdamW's (train_multimodal_fusion.py:433-439 saves it as `optimizer_state_dict`):
        per-parameter `step` as a CPU fp32 scalar tensor, `exp_avg`, `exp_avg_sq`, so a stock AdamW can load it."""
        sd = super().state_dict()
        sd["state"] = {k: dict(v) for k, v in sd["state"].items()}      # the base class hands out the live per-param dicts
        host_steps = {k: v.detach().float().cpu() for k, v in self._steps.items()}
        for st in sd["state"].values():
            if "step" in st and torch.is_tensor(st["step"]):
                dev = st["step"].device
                st["step"] = host_steps.get((dev.type, dev.index), st["step"].detach().float().cpu()).clone()
        return sd

    def load_state_dict(self, state_dict):
        """Accepts checkpoints of this class and of torch.optim.AdamW (any device, any memory layout of the moments):
        `step` becomes the shared device counter, the moments are re-laid-out like their parameter (conv weights are
        channels_last here, NCHW in a reference checkpoint), and every cached pointer table is dropped."""
        super().load_state_dict(state_dict)
        self._tables.clear()
        self._steps = {}
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state.get(p)
                if not st:
                    continue
                key = (p.device.type, p.device.index)
                step = st.get("step", 0.0)
                step = float(step.item()) if torch.is_tensor(step) else float(step)
                if key not in self._steps:
                    self._steps[key] = torch.full((), step, dtype=torch.float32, device=p.device)
                elif float(self._steps[key].item()) != step:
                    raise RuntimeError("FusedAdamW keeps one step counter per device; the checkpoint holds different "
                                       "step counts (%g vs %g)" % (float(self._steps[key].item()), step))
                st["step"] = self._steps[key]
                for name in ("exp_avg", "exp_avg_sq"):
                    src = st.get(name)
                    if src is None:
                        src = torch.zeros((), dtype=torch.float32).expand(p.shape)
                    dst = torch.empty_like(p, dtype=torch.float32, memory_format=torch.preserve_format)
                    dst.copy_(src.to(p.device))
                    st[name] = dst

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        L = _lib.load()
        st = torch.cuda.current_stream().cuda_stream
        ticked = set()         # the shared counter advances once per step() however many groups read it
        ops.PARAM_EPOCH[0] += 1        # parameters change below without torch's version counters noticing (folded-BN caches)
        for gi, group in enumerate(self.param_groups):
            if group.get("amsgrad") or group.get("maximize") or group.get("decoupled_weight_decay", True) is False:
                raise RuntimeError("FusedAdamW implements AdamW as the reference configures it: no amsgrad / maximize, decoupled weight decay")
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
            if len(set((p.device.type, p.device.index) for p in params)) > 1:
                raise RuntimeError("FusedAdamW: the parameters of one group must live on one device")
            # one launch per gradient bucket when a bucketer is attached (each waits for its own all-reduce), else one launch
            parts = [(None, params, grads)]
            if self._bucketer is not None:
                by_bucket = {}
                for p, g in zip(params, grads):
                    by_bucket.setdefault(self._bucketer.bucket_of(p), ([], []))
                    by_bucket[self._bucketer.bucket_of(p)][0].append(p)
                    by_bucket[self._bucketer.bucket_of(p)][1].append(g)
                parts = [(bi, ps, gs) for bi, (ps, gs) in sorted(by_bucket.items(), key=lambda kv: (kv[0] is None, kv[0]))]
            b1, b2 = group["betas"]
            for bi, ps, gs in parts:
                cached = self._table_for((gi, bi), group, ps, gs, g_bf16)
                if cached["step"].data_ptr() not in ticked:
                    ticked.add(cached["step"].data_ptr())
                    check(L.dfu_adamw_tick(cached["step"].data_ptr(), st), "adamw_tick")
                    ops._count()
                if bi is not None:
                    self._bucketer.wait_bucket(bi)
                check(L.dfu_adamw(cached["table"].data_ptr(), cached["n"], cached["ct"].data_ptr(), cached["co"].data_ptr(),
                                  cached["n_chunks"], CHUNK, cached["step"].data_ptr(), group["lr"], b1, b2, group["eps"],
                                  group["weight_decay"], scale, 1 if cached["g_bf16"] else 0, st), "adamw")
                ops._count()
                fresh = set(id(p) for p in cached["shadowed"])
                for p in ps:
                    if id(p) in fresh:
                        mark_shadow_fresh(p)
                    elif getattr(p, "_dfu_shadow", None) is not None:
                        p._dfu_shadow = (None, p._dfu_shadow[1])     # padded shadow: recast on next use
        return loss
