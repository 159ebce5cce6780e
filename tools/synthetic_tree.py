"""Synthetic DFU data tree in the directory layout the reference drivers hard-wire
(notebooks/train_multimodal_fusion.py:46-49, extended_metrics.py:279-283):

    <home>/CompVis/Dataset/data/{rgb,thermal}/{train,val,test}/{healthy,ulcer}/*.png

Images are seeded noise of varying sizes (so the scripts' Resize((224,224)) really resamples) with a central patch that
is brighter for `ulcer` and darker for `healthy`, every file distinct (the scripts run a SHA-256 leakage check across
splits, train_multimodal_fusion.py:230-265).  Harness code: used by tests/test_gpu_reference_drivers.py to run the
reference's own scripts against the B200 drop-in; not part of the product path."""
import os

import numpy as np
from PIL import Image

DEFAULT_COUNTS = {"train": 6, "val": 3, "test": 4}          # images per class, per modality


def make_tree(home, counts=None, seed=0, sizes=((240, 200), (224, 224), (300, 260), (180, 232))):
    counts = dict(DEFAULT_COUNTS, **(counts or {}))
    rng = np.random.RandomState(seed)
    data = os.path.join(home, "CompVis", "Dataset", "data")
    written = []
    k = 0
    for modality in ("rgb", "thermal"):
        for split, n in counts.items():
            for label, cls in enumerate(("healthy", "ulcer")):
                d = os.path.join(data, modality, split, cls)
                os.makedirs(d, exist_ok=True)
                # the thermal set is one image larger per class: the scripts pair by cycling the shorter list
                for i in range(n + (1 if modality == "thermal" else 0)):
                    w, h = sizes[k % len(sizes)]
                    k += 1
                    img = rng.randint(40, 216, size=(h, w, 3)).astype(np.int16)
                    cy, cx, r = h // 2, w // 2, min(h, w) // 4
                    img[cy - r:cy + r, cx - r:cx + r, :] += (40 if label == 1 else -40)
                    if modality == "thermal":
                        img[..., 1] = img[..., 0]
                        img[..., 2] = img[..., 0]                  # grey-scale thermogram stored as RGB
                    path = os.path.join(d, "%s_%s_%s_%03d.png" % (modality, split, cls, i))
                    Image.fromarray(np.clip(img, 0, 255).astype(np.uint8)).save(path)
                    written.append(path)
    return data, written


if __name__ == "__main__":
    import sys
    root, files = make_tree(sys.argv[1] if len(sys.argv) > 1 else "/tmp/dfu_home")
    print(root, len(files), "files")
