"""Host -> device input staging for the train / eval loops.

The reference moves every batch with three synchronous `.to(DEVICE)` calls inside the hot loop
(notebooks/train_multimodal_fusion.py:369-371).  `DevicePrefetcher` keeps the same loop shape but issues the copies of
batch i+1 from pinned host memory on a side stream while step i computes, so the H2D transfer (38.5 MB of uint8 image
pairs at B=128) leaves the critical path.  Device buffers are double-buffered and reused, which also keeps the
addresses stable for CUDA-graph replays of the step."""
import torch


class DevicePrefetcher:
    """Iterates over `batches` (an iterable of tuples of pinned CPU tensors) yielding tuples of device tensors."""

    def __init__(self, batches, device, depth=2):
        self.batches = batches
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DevicePrefetcher stages batches for the CUDA path; got device %s" % (device,))
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth = depth
        self.slots = [None] * depth
        self.ready = [None] * depth
        self.consumed = [None] * depth

    def _stage(self, slot, batch):
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.stream):
            if self.consumed[slot] is not None:
                self.stream.wait_event(self.consumed[slot])     # the step that read this slot has finished
            if self.slots[slot] is None or any(b.shape != s.shape or b.dtype != s.dtype
                                               for b, s in zip(batch, self.slots[slot])):
                self.slots[slot] = tuple(torch.empty(b.shape, dtype=b.dtype, device=self.device) for b in batch)
            for dst, src in zip(self.slots[slot], batch):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.ready[slot] = ev
        del main

    def __iter__(self):
        it = iter(self.batches)
        pending = []
        slot = 0
        for _ in range(self.depth - 1):
            try:
                self._stage(slot, next(it))
            except StopIteration:
                break
            pending.append(slot)
            slot = (slot + 1) % self.depth
        while pending:
            try:
                self._stage(slot, next(it))
                pending.append(slot)
                slot = (slot + 1) % self.depth
            except StopIteration:
                pass
            cur = pending.pop(0)
            main = torch.cuda.current_stream(self.device)
            main.wait_event(self.ready[cur])
            yield self.slots[cur]
            done = torch.cuda.Event()
            done.record(main)
            self.consumed[cur] = done


# ======================================================================================================================
# Ingest: the reference's MultimodalDataset + WeightedRandomSampler + DataLoader
# (notebooks/train_multimodal_fusion.py:60-167, 268-280) with everything after the JPEG/PNG decode moved to the GPU.
# ======================================================================================================================
import os          # noqa: E402
import random      # noqa: E402
from concurrent.futures import ThreadPoolExecutor   # noqa: E402

IMAGE_EXTS = {".jpg", ".jpeg", ".png", ".bmp", ".tif", ".tiff"}


def _list_images(d):
    out = []
    if os.path.isdir(d):
        for root, _, files in os.walk(d):
            out += [os.path.join(root, f) for f in files if os.path.splitext(f)[1].lower() in IMAGE_EXTS]
    return sorted(out)


class PairedFolderDataset:
    """The reference's label-matched pseudo-pairing (train_multimodal_fusion.py:108-140): within each class the shorter
    modality list is cycled against the longer one (`i % len`), healthy pairs first, then ulcer pairs, then the list is
    shuffled with `random.shuffle` -- pass the `random.Random` the run is seeded with to reproduce the reference's order.
    Items are RAW decoded uint8 images (any size): resize / augmentation / normalisation happen on the device (GpuIngest)."""

    def __init__(self, rgb_dir, thermal_dir, split="train", shuffle_rng=None):
        self.pairs = []
        for label, cls in enumerate(("healthy", "ulcer")):
            rgb = _list_images(os.path.join(rgb_dir, split, cls))
            th = _list_images(os.path.join(thermal_dir, split, cls))
            if rgb and th:
                for i in range(max(len(rgb), len(th))):
                    self.pairs.append((rgb[i % len(rgb)], th[i % len(th)], label))
        (shuffle_rng or random).shuffle(self.pairs)
        self.labels = [p[2] for p in self.pairs]

    def __len__(self):
        return len(self.pairs)

    def load(self, idx):
        """(rgb uint8 [H,W,3], thermal uint8 [H',W',3], label) -- Image.open(path).convert('RGB') as the reference (:155-156)."""
        import numpy as np
        from PIL import Image
        rp, tp, label = self.pairs[idx]
        return (np.asarray(Image.open(rp).convert("RGB")), np.asarray(Image.open(tp).convert("RGB")), label)


def class_weights(labels):
    """CrossEntropyLoss weights of the reference: total / count per class (train_multimodal_fusion.py:342-346)."""
    n0, n1 = labels.count(0), labels.count(1)
    tot = max(n0 + n1, 1)
    return torch.tensor([tot / n0 if n0 else 0.0, tot / n1 if n1 else 0.0], dtype=torch.float32)


def weighted_sample_indices(labels, generator=None):
    """One epoch of WeightedRandomSampler(1 / class_count[label], num_samples=len, replacement=True) (:268-273)."""
    counts = [labels.count(0), labels.count(1)]
    w = torch.tensor([1.0 / counts[y] if counts[y] else 0.0 for y in labels], dtype=torch.double)
    return torch.multinomial(w, len(labels), replacement=True, generator=generator).tolist()


class GpuIngest:
    """Iterates over a PairedFolderDataset in batches and yields (rgb, thermal, labels) ready for MultimodalFusionModel:
    rgb = normalised bf16 (logical (B,3,224,224), NHWC memory), thermal = normalised bf16 ViT patch matrix [B*196, 768].

    Host side: PIL decode in a thread pool, the raw uint8 pixels of a batch packed into ONE pinned buffer, one H2D copy on a
    side stream while the previous batch trains.  Device side: Resize((224,224)) bit-identical to the reference's
    (ops.resize_u8, one launch per distinct source size), then -- train=True -- the reference's random augmentation
    (augment.draw_params / augment_u8) and ToTensor + Normalize fused into the same kernels."""

    def __init__(self, dataset, batch_size, device, train=False, sampler_generator=None, aug_rng=None, workers=4):
        self.ds, self.bs, self.train = dataset, batch_size, train
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GpuIngest feeds the CUDA path (no CPU fallback); got %s" % (device,))
        self.gen, self.aug_rng, self.workers = sampler_generator, aug_rng or random.Random(0), workers
        self.stream = torch.cuda.Stream(device=self.device)

    def __len__(self):
        return (len(self.ds) + self.bs - 1) // self.bs

    def _order(self):
        if self.train:
            return weighted_sample_indices(self.ds.labels, self.gen)
        return list(range(len(self.ds)))

    def _stage(self, idxs, pool):
        """Decode, order the batch by source size (so that images of one size form contiguous runs: one resize launch per
        run), pack each modality into one pinned buffer, copy on the side stream."""
        import numpy as np
        items = list(pool.map(self.ds.load, idxs))
        order = sorted(range(len(items)), key=lambda i: (items[i][0].shape[:2], items[i][1].shape[:2]))
        items = [items[i] for i in order]                                      # the order inside a batch carries no meaning
        batch_indices = [idxs[i] for i in order]
        labels = torch.tensor([it[2] for it in items], dtype=torch.int64).pin_memory()
        staged = []
        for mod in (0, 1):
            imgs = [np.ascontiguousarray(it[mod]) for it in items]
            flat = torch.empty(sum(im.size for im in imgs), dtype=torch.uint8).pin_memory()
            runs, off = [], 0
            for im in imgs:
                flat[off:off + im.size] = torch.from_numpy(im.reshape(-1))
                hw = tuple(im.shape[:2])
                if runs and runs[-1][0] == hw:
                    runs[-1][2] += 1
                else:
                    runs.append([hw, off, 1])
                off += im.size
            with torch.cuda.stream(self.stream):
                dev = flat.to(self.device, non_blocking=True)
            staged.append((dev, runs, flat))                                   # `flat` stays alive until the copy has run
        with torch.cuda.stream(self.stream):
            dlabels = labels.to(self.device, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return staged, dlabels, ready, (labels, batch_indices)

    def _finish(self, staged, dlabels, ready, keep):
        from . import augment, ops
        self.last_indices = keep[1]           # dataset indices of the batch being yielded, in batch order
        torch.cuda.current_stream().wait_event(ready)
        B = dlabels.shape[0]
        outs = []
        for (dev, runs, _flat), modality, kind, mode in ((staged[0], "rgb", "imagenet", "nhwc"), (staged[1], "thermal", "half", "patches")):
            dev.record_stream(torch.cuda.current_stream())
            u8 = torch.empty((B, 224, 224, 3), dtype=torch.uint8, device=self.device)
            pos = 0
            for (h, w), off, n in runs:
                src = dev[off:off + n * h * w * 3].view(n, h, w, 3)
                if (h, w) == (224, 224):
                    u8[pos:pos + n].copy_(src)                                 # a device-to-device memcpy
                else:
                    ops.resize_u8(src, (224, 224), "u8", out=u8[pos:pos + n])
                pos += n
            if self.train:
                x = augment.augment_u8(u8, augment.draw_params(B, modality, rng=self.aug_rng), mode, kind)
            else:
                x = ops.preprocess_u8_nhwc(u8, kind) if mode == "nhwc" else ops.preprocess_u8_patches(u8, kind)
            outs.append(x)
        dlabels.record_stream(torch.cuda.current_stream())
        return outs[0], outs[1], dlabels

    def __iter__(self):
        order = self._order()
        batches = [order[i:i + self.bs] for i in range(0, len(order), self.bs)]
        with ThreadPoolExecutor(max_workers=self.workers) as pool:
            nxt = self._stage(batches[0], pool) if batches else None
            for k in range(len(batches)):
                cur = nxt
                nxt = self._stage(batches[k + 1], pool) if k + 1 < len(batches) else None      # overlaps the consumer's step
                yield self._finish(*cur)
