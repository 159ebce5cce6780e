"""One training step of the fusion model as the reference's hot loop runs it
(notebooks/train_multimodal_fusion.py:368-388: zero_grad -> forward -> CrossEntropyLoss(weight) -> backward ->
AdamW.step), optionally captured into ONE CUDA graph so the ~2500 kernel launches of a step replay without any
Python or launch-queue stalls, and optionally data-parallel (ddp.GradBucketer).  The loss stays on the device;
the caller decides when to read it (the reference's per-step .item() sync, :383, is what this removes)."""
import torch

from . import ops
from .functional import cross_entropy
from .resnet import refresh_folded


class TrainStep:
    """Captured CUDA graphs are keyed by the batch's shapes and dtypes: the reference DataLoaders do not drop the last,
    smaller batch of an epoch (train_multimodal_fusion.py:272-275), so a new shape gets its own graph (up to
    `max_graphs`) and anything beyond that runs eagerly -- never a silent broadcast into the first batch's buffers."""

    def __init__(self, model, optimizer, class_weights=None, use_graph=True, bucketer=None, warmup=3, max_graphs=3):
        self.model, self.optimizer, self.bucketer = model, optimizer, bucketer
        self.class_weights = class_weights
        self.use_graph = use_graph
        self.warmup = warmup
        self.max_graphs = max_graphs
        self.graphs = {}            # (shape, dtype) signature of the batch -> (graph, static inputs, static outputs, launches)
        self.graph = None           # the most recently replayed graph (None while running eagerly)
        self.calls = 0
        self.launches_per_step = None

    # -- the step itself (eager) -------------------------------------------------------------------------
    def _step(self, *args):
        *inputs, labels = args
        self.optimizer.zero_grad(set_to_none=True)
        ops.ZERO_ARENA.begin_step(labels.device)   # one dfu_zero launch clears every accumulator of the step
        if self.bucketer is not None:
            self.bucketer.begin_step()
        logits = self.model(*inputs)
        loss = cross_entropy(logits, labels, self.class_weights)
        loss.backward()
        if self.bucketer is not None:
            if getattr(self.optimizer, "_bucketer", None) is self.bucketer:
                self.bucketer.flush()            # FusedAdamW waits bucket by bucket (optim.py), not for the whole exchange
            else:
                self.bucketer.finish_step()
        self.optimizer.step()
        ops.ZERO_ARENA.end_step()
        return loss.detach(), logits.detach()

    def _eager(self, *args):
        n0 = ops.LAUNCHES
        out = self._step(*args)
        self.launches_per_step = ops.LAUNCHES - n0
        return out

    @staticmethod
    def _key(*tensors):
        return tuple((tuple(t.shape), t.dtype) for t in tensors)

    def _capture(self, key, *args):
        static = tuple(torch.empty_like(t) for t in args)
        for s, t in zip(static, args):
            s.copy_(t)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = ops.LAUNCHES
        with torch.cuda.graph(graph):
            static_out = self._step(*static)
        self.graphs[key] = (graph, static, static_out, ops.LAUNCHES - n0)
        return self.graphs[key]

    def __call__(self, *args):
        """ts(rgb, thermal, labels) for the fusion model, ts(images, labels) for a single-modality model.  Inputs are
        device tensors (uint8 NHWC or fp32 NCHW), labels int64.  Returns (loss, logits) on the device."""
        self.model.train()
        if len(args) < 2:
            raise RuntimeError("TrainStep: expected (*model_inputs, labels)")
        B = args[-1].shape[0]
        # image batches are (B, ...) 4-D tensors; an already unfolded ViT patch matrix is [B*196, 768]
        if any((t.shape[0] != B) if t.dim() != 2 else (t.shape[0] not in (B, B * 196)) for t in args):
            raise RuntimeError("TrainStep: batch sizes differ: %s" % ([tuple(t.shape) for t in args],))
        if not self.use_graph:
            self.graph = None
            return self._eager(*args)
        key = self._key(*args)
        entry = self.graphs.get(key)
        if entry is None:
            if self.calls < self.warmup or len(self.graphs) >= self.max_graphs:
                # eager: the warm-up calls allocate optimizer state and weight shadows and fill the allocator pools;
                # later, batch shapes beyond the graph budget simply run uncaptured
                self.calls += 1
                self.graph = None
                return self._eager(*args)
            entry = self._capture(key, *args)
        graph, static, static_out, launches = entry
        for s, t in zip(static, args):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        graph.replay()
        ops._count(launches)          # replayed kernels count like launched ones (bench.py's gpu_launches)
        self.graph, self.launches_per_step = graph, launches
        return static_out


class EvalStep:
    """Inference forward (extended_metrics.py:623-642) with the softmax/argmax epilogue kernel; one captured CUDA graph
    per batch shape (the last batch of a loader is smaller)."""

    def __init__(self, model, use_graph=True, max_graphs=4):
        self.model, self.use_graph, self.max_graphs = model, use_graph, max_graphs
        self.graphs = {}
        self.graph = None
        self.launches_per_step = None

    @torch.no_grad()
    def _fwd(self, *inputs):
        logits = self.model(*inputs)
        prob1, pred = ops.predict(logits)
        return logits, prob1, pred

    def __call__(self, *inputs):
        """step(rgb, thermal) or step(images): (logits [B,C] fp32, softmax[:,1] [B] fp32, argmax [B] int64), on the device."""
        self.model.eval()
        B = inputs[0].shape[0]
        if any((t.shape[0] != B) if t.dim() != 2 else (t.shape[0] not in (B, B * 196)) for t in inputs):
            raise RuntimeError("EvalStep: batch sizes differ: %s" % ([tuple(t.shape) for t in inputs],))
        refresh_folded(self.model)       # folded conv/BatchNorm weights are inputs of the captured graphs, not part of them
        key = TrainStep._key(*inputs)
        entry = self.graphs.get(key) if self.use_graph else None
        if entry is None:
            if not self.use_graph or len(self.graphs) >= self.max_graphs:
                self.graph = None
                return self._fwd(*inputs)
            self._fwd(*inputs)           # warm-up (weight shadows, allocator)
            static = tuple(torch.empty_like(t) for t in inputs)
            for s, t in zip(static, inputs):
                s.copy_(t)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = ops.LAUNCHES
            with torch.cuda.graph(graph):
                static_out = self._fwd(*static)
            entry = self.graphs[key] = (graph, static, static_out, ops.LAUNCHES - n0)
        graph, static, static_out, launches = entry
        for s, t in zip(static, inputs):
            if s.data_ptr() != t.data_ptr():
                s.copy_(t, non_blocking=True)
        graph.replay()
        ops._count(launches)
        self.graph, self.launches_per_step = graph, launches
        return static_out
