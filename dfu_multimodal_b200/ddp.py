"""Data parallelism for the fusion train step (BASELINE config 4; the reference is single-process).

One process per GPU.  Gradients are cast fp32 -> bf16 into ~25 MB flat buckets as autograd produces them
(post-accumulate-grad hooks, i.e. in reverse layer order) and each full bucket is all-reduced with NCCL on a
side stream while backward keeps running; FusedAdamW then reads the averaged bf16 buckets directly.  BatchNorm
statistics stay per rank (the reference has no SyncBN)."""
import torch
import torch.distributed as dist

from . import ops

BUCKET_BYTES = 25 * 1024 * 1024


CAST_CHUNK = 65536


def _default_cast(src_f32, dst_bf16):
    """fp32 gradient (dense, any dim order) -> bf16 bucket slice in storage order, with our cast kernel."""
    ops.cast_f32_bf16(src_f32, 1, src_f32.numel(), src_f32.numel(), out=dst_bf16)


class _BucketCaster:
    """All gradients of a bucket -> its bf16 wire buffer with ONE dfu_cast_multi_f32_bf16 launch (317 per-parameter cast
    launches per step otherwise).  The device pointer tables are cached per (bucket, gradient addresses): inside a captured
    CUDA graph the addresses are fixed, so each table is built once."""

    def __init__(self):
        self.tables = {}

    def __call__(self, bi, grads, dsts):
        import ctypes
        from . import _lib
        from ._lib import CastTensor, check
        key = (bi, tuple(g.data_ptr() for g in grads))
        ent = self.tables.get(key)
        if ent is None:
            n = len(grads)
            arr = (CastTensor * n)()
            ct, co = [], []
            for i, (g, d) in enumerate(zip(grads, dsts)):
                arr[i].src, arr[i].dst, arr[i].numel = g.data_ptr(), d.data_ptr(), g.numel()
                for c in range((g.numel() + CAST_CHUNK - 1) // CAST_CHUNK):
                    ct.append(i)
                    co.append(c)
            host = [torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).pin_memory(),
                    torch.tensor(ct, dtype=torch.int32).pin_memory(), torch.tensor(co, dtype=torch.int32).pin_memory()]
            dev = [h.to(grads[0].device, non_blocking=True) for h in host]       # memcpy nodes under graph capture
            ent = dict(host=host, dev=dev, n=n, n_chunks=len(ct), captured=torch.cuda.is_current_stream_capturing())
            eager = [k for k, e in self.tables.items() if not e["captured"]]
            for k in eager[:max(0, len(eager) - 64)]:
                del self.tables[k]
            self.tables[key] = ent
        t, c, o = ent["dev"]
        check(_lib.load().dfu_cast_multi_f32_bf16(t.data_ptr(), ent["n"], c.data_ptr(), o.data_ptr(), ent["n_chunks"], CAST_CHUNK,
                                                  torch.cuda.current_stream().cuda_stream), "cast_multi_f32_bf16")
        ops._count()


class GradBucketer:
    def __init__(self, params, process_group=None, bucket_bytes=BUCKET_BYTES, cast_fn=None, average=True,
                 comm_stream=None, overlap=True):
        """overlap=True: a bucket's all-reduce starts as soon as its last gradient is there, beside the rest of backward.
        overlap=False: the all-reduces start together when backward is done (flush / finish_step) and only overlap the
        bucket-by-bucket optimizer update -- no NCCL CTAs compete with the persistent GEMMs of backward for SMs."""
        self.overlap = overlap
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.cast_fn = cast_fn or _default_cast
        # CUDA + our own cast: the gradients of a bucket are cast together when the bucket is complete
        self.bucket_caster = _BucketCaster() if cast_fn is None else None
        self.average = average
        params = [p for p in params if p.requires_grad]
        self.params = params
        # reverse registration order ~ the order backward produces gradients (head first, stems last)
        order = list(reversed(params))
        self.buckets = []
        cur, cur_elems = [], 0
        limit = bucket_bytes // 2
        for p in order:
            n = (p.numel() + 7) // 8 * 8          # keep every slice 16-byte aligned
            if cur and cur_elems + n > limit:
                self.buckets.append(cur)
                cur, cur_elems = [], 0
            cur.append(p)
            cur_elems += n
        if cur:
            self.buckets.append(cur)
        self.flat = []
        self.slot = {}
        for bi, plist in enumerate(self.buckets):
            total = sum((p.numel() + 7) // 8 * 8 for p in plist)
            flat = torch.zeros(total, dtype=torch.bfloat16, device=plist[0].device)
            self.flat.append(flat)
            off = 0
            for p in plist:
                self.slot[p] = (bi, off, p.numel())
                off += (p.numel() + 7) // 8 * 8
        self.pending = [0] * len(self.buckets)
        self.producers = [set() for _ in self.buckets]
        self.got = set()          # parameters that produced a gradient since begin_step()
        self.works = []
        self.events = [None] * len(self.buckets)     # bucket -> CUDA event recorded behind its all-reduce
        self.launched = [False] * len(self.buckets)
        self.ready = [[] for _ in self.buckets]      # (gradient, wire slice) pairs waiting for the bucket's one cast launch
        self.is_cuda = params[0].is_cuda
        self.comm_stream = comm_stream if comm_stream is not None else (torch.cuda.Stream() if self.is_cuda else None)
        self.handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in params]

    # ---------------------------------------------------------------------------------------------------
    def begin_step(self):
        self.pending = [len(b) for b in self.buckets]
        self.producers = [set() for _ in self.buckets]
        self.got = set()
        self.works = []
        self.events = [None] * len(self.buckets)
        self.launched = [False] * len(self.buckets)
        self.ready = [[] for _ in self.buckets]

    def _on_grad(self, p):
        bi, off, n = self.slot[p]
        self.got.add(p)
        g = p.grad
        same_order = all(gs == ps for gs, ps, k in zip(g.stride(), p.stride(), p.shape) if k > 1)   # size-1 dims carry any stride
        if not same_order:
            g = g.contiguous() if p.is_contiguous() else torch.empty_like(p).copy_(g)
        batched = self.bucket_caster is not None and self.is_cuda and g.dtype == torch.float32
        if batched:
            self.ready[bi].append((g, self.flat[bi][off:off + n]))
        else:
            self.cast_fn(g, self.flat[bi][off:off + n])
        if self.is_cuda:
            # the two model branches may run backward on different streams; remember every producer stream
            self.producers[bi].add(torch.cuda.current_stream())
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._cast_bucket(bi)
            if self.overlap:
                self._launch(bi)

    def _cast_bucket(self, bi):
        """One launch puts every gradient of the bucket on the wire buffer (on the current stream, after the other producers)."""
        if not self.ready[bi]:
            return
        cur = torch.cuda.current_stream()
        for s in self.producers[bi]:
            if s != cur:
                cur.wait_stream(s)
        grads, dsts = zip(*self.ready[bi])
        self.bucket_caster(bi, grads, dsts)
        self.ready[bi] = []

    def _launch(self, bi):
        self.launched[bi] = True
        if self.world == 1:
            return
        flat = self.flat[bi]
        op = dist.ReduceOp.AVG if (self.average and self.is_cuda) else dist.ReduceOp.SUM
        if self.is_cuda:
            self.comm_stream.wait_stream(torch.cuda.current_stream())
            for s in self.producers[bi]:
                self.comm_stream.wait_stream(s)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(flat, op=op, group=self.pg)
                ev = torch.cuda.Event()
                ev.record(self.comm_stream)
                self.events[bi] = ev
        else:
            self.works.append(dist.all_reduce(flat, op=op, group=self.pg, async_op=True))

    def bucket_of(self, p):
        return self.slot[p][0] if p in self.slot else None

    def wait_bucket(self, bi):
        """The current stream waits for bucket `bi`'s all-reduce only (FusedAdamW updates bucket by bucket, so the update of
        the early buckets runs while the last all-reduces are still on the wire)."""
        if self.is_cuda and self.events[bi] is not None:
            torch.cuda.current_stream().wait_event(self.events[bi])

    def flush(self):
        """Launch the all-reduce of every bucket that is still waiting for gradients that will not come (frozen or unused
        parameters); does NOT wait for any collective."""
        self._flush_pending()

    def finish_step(self):
        """All buckets reduced and visible to the current stream."""
        self._flush_pending()
        if self.is_cuda:
            if self.world > 1:
                torch.cuda.current_stream().wait_stream(self.comm_stream)
        else:
            for w in self.works:
                w.wait()
            self.works = []

    def _flush_pending(self):
        for bi, left in enumerate(self.pending):
            if self.launched[bi]:
                continue
            if left > 0:
                # some parameters of this bucket received no gradient this step (frozen or conditionally used): their
                # slices still hold LAST step's reduced gradient -- clear them so the all-reduce moves zeros, and
                # grad_source() reports them as gradient-less so the optimizer skips them like torch.optim.AdamW does
                for p in self.buckets[bi]:
                    if p not in self.got:
                        _, off, n = self.slot[p]
                        self.flat[bi][off:off + n].zero_()
                self.pending[bi] = 0
                if self.is_cuda:
                    self._cast_bucket(bi)
            self._launch(bi)                         # (deferred mode: complete buckets go out here too, in bucket order)

    def grad_source(self, p):
        """For FusedAdamW.set_grad_source: (bf16 grad slice, is_bf16, scale)."""
        if p not in self.slot or p not in self.got:
            return None
        bi, off, n = self.slot[p]
        scale = 1.0
        if self.world > 1 and self.average and not self.is_cuda:
            scale = 1.0 / self.world     # gloo has no AVG
        return self.flat[bi][off:off + n], True, scale

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []


def broadcast_parameters(model, src=0, group=None):
    """Identical replicas at start (rank-0 weights and BN buffers)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src=src, group=group)
        if hasattr(t, "_dfu_shadow"):
            t._dfu_shadow = (None, t._dfu_shadow[1])
