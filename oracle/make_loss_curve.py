"""ORACLE / TEST INFRASTRUCTURE.  Generates tests/golden/loss_curve_200_seed1.json: the loss of the reference hot loop
(notebooks/train_multimodal_fusion.py:368-388, restated in oracle/reference_model.py: fp32 CPU, CrossEntropyLoss with the
276-pair class weights, AdamW lr 1e-4 wd 1e-4) over 200 steps of fresh synthetic batches of 8 pairs, head dropout off.
tests/test_gpu_model.py::test_loss_curve_200_steps_against_golden replays the same 200 batches on the CUDA path.
Usage (CPU, ~10 min): python oracle/make_loss_curve.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_model as R  # noqa: E402

STEPS, B, SEED, BATCH_SEED0 = 200, 8, 1, 500


def main():
    oracle = R.seeded_init_(R.OracleFusion(), SEED)
    for m in oracle.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    w = torch.tensor([276 / 130, 276 / 146])
    losses = []
    opt = torch.optim.AdamW(oracle.parameters(), lr=1e-4, weight_decay=1e-4)
    crit = torch.nn.CrossEntropyLoss(weight=w)
    oracle.train()
    for i in range(STEPS):
        a, b, y = R.synthetic_batch(B, BATCH_SEED0 + i)
        opt.zero_grad()
        loss = crit(oracle(R.preprocess(a, "imagenet"), R.preprocess(b, "half")), y)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
        if i % 20 == 0:
            print(i, losses[-1], flush=True)
    out = os.path.join(ROOT, "tests", "golden", "loss_curve_200_seed1.json")
    with open(out, "w") as f:
        json.dump({"steps": STEPS, "batch": B, "init_seed": SEED, "batch_seed0": BATCH_SEED0, "oracle": losses}, f)
    print("wrote", out)


if __name__ == "__main__":
    main()
