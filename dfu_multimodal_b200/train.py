"""One training step of the fusion model as the reference's hot loop runs it
(notebooks/train_multimodal_fusion.py:368-388: zero_grad -> forward -> CrossEntropyLoss(weight) -> backward ->
AdamW.step), optionally captured into ONE CUDA graph so the ~2500 kernel launches of a step replay without any
Python or launch-queue stalls, and optionally data-parallel (ddp.GradBucketer).  The loss stays on the device;
the caller decides when to read it (the reference's per-step .item() sync, :383, is what this removes)."""
import torch

from . import ops
from .functional import cross_entropy


class TrainStep:
    def __init__(self, model, optimizer, class_weights=None, use_graph=True, bucketer=None, warmup=3):
        self.model, self.optimizer, self.bucketer = model, optimizer, bucketer
        self.class_weights = class_weights
        self.use_graph = use_graph
        self.warmup = warmup
        self.graph = None
        self.static = None
        self.calls = 0
        self.launches_per_step = None

    # -- the step itself (eager) -------------------------------------------------------------------------
    def _step(self, rgb, thermal, labels):
        self.optimizer.zero_grad(set_to_none=True)
        ops.ZERO_ARENA.begin_step(rgb.device)      # one dfu_zero launch clears every accumulator of the step
        if self.bucketer is not None:
            self.bucketer.begin_step()
        ops.DROPOUT.tick(rgb.device)
        logits = self.model(rgb, thermal)
        loss = cross_entropy(logits, labels, self.class_weights)
        loss.backward()
        if self.bucketer is not None:
            self.bucketer.finish_step()
        self.optimizer.step()
        ops.ZERO_ARENA.end_step()
        return loss.detach(), logits.detach()

    def _capture(self, rgb, thermal, labels):
        self.static = (torch.empty_like(rgb), torch.empty_like(thermal), torch.empty_like(labels))
        for s, t in zip(self.static, (rgb, thermal, labels)):
            s.copy_(t)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.LAUNCHES
        with torch.cuda.graph(self.graph):
            self.static_out = self._step(*self.static)
        self.launches_per_step = ops.LAUNCHES - n0

    def __call__(self, rgb, thermal, labels):
        """rgb/thermal: device tensors (uint8 NHWC or fp32 NCHW); labels int64. Returns (loss, logits) on device."""
        self.model.train()
        if not self.use_graph:
            n0 = ops.LAUNCHES
            out = self._step(rgb, thermal, labels)
            self.launches_per_step = ops.LAUNCHES - n0
            return out
        if self.graph is None:
            if self.calls < self.warmup:
                # eager warm-up: allocates optimizer state and weight shadows, fills the allocator pools
                self.calls += 1
                n0 = ops.LAUNCHES
                out = self._step(rgb, thermal, labels)
                self.launches_per_step = ops.LAUNCHES - n0
                return out
            self._capture(rgb, thermal, labels)
        for s, t in zip(self.static, (rgb, thermal, labels)):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_out


class EvalStep:
    """Inference forward (extended_metrics.py:623-642) with the softmax/argmax epilogue kernel; graph-captured."""

    def __init__(self, model, use_graph=True):
        self.model, self.use_graph = model, use_graph
        self.graph = None

    @torch.no_grad()
    def _fwd(self, rgb, thermal):
        logits = self.model(rgb, thermal)
        prob1, pred = ops.predict(logits)
        return logits, prob1, pred

    def __call__(self, rgb, thermal):
        self.model.eval()
        if not self.use_graph:
            return self._fwd(rgb, thermal)
        if self.graph is None or self.static[0].shape != rgb.shape:
            self._fwd(rgb, thermal)      # warm-up (weight shadows, allocator)
            self.static = (torch.empty_like(rgb), torch.empty_like(thermal))
            for s, t in zip(self.static, (rgb, thermal)):
                s.copy_(t)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_out = self._fwd(*self.static)
        for s, t in zip(self.static, (rgb, thermal)):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_out
