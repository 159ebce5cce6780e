"""ORACLE / TEST INFRASTRUCTURE.  Generates the committed loss-curve fixtures of the reference hot loop
(notebooks/train_multimodal_fusion.py:368-388, restated in oracle/reference_model.py: fp32 CPU, CrossEntropyLoss with
class weights, AdamW lr 1e-4 wd 1e-4), head dropout off:

  learning (default)  tests/golden/loss_curve_learning_200_seed1.json -- a FIXED set of 64 synthetic pairs cycled as
                      8 batches of 8 for 200 steps: the network memorises them, the loss falls from ln 2 to < 1e-3
                      (below 0.1 after ~26 steps), so a broken backward or optimizer cannot track it.
  random              tests/golden/loss_curve_200_seed1.json -- 200 fresh random-label batches (the bench regime; the
                      loss stays near ln 2, kept as a stability check).

tests/test_gpu_model.py replays the same batches on the CUDA path.  Usage (CPU, ~10 min each):
  python oracle/make_loss_curve.py [learning|random]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_model as R  # noqa: E402

STEPS, B, SEED, BATCH_SEED0 = 200, 8, 1, 500
SET_SIZE, SET_SEED = 64, 4242


def run(batch_of_step, weight, out_name, extra):
    oracle = R.seeded_init_(R.OracleFusion(), SEED)
    for m in oracle.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    losses = []
    opt = torch.optim.AdamW(oracle.parameters(), lr=1e-4, weight_decay=1e-4)
    crit = torch.nn.CrossEntropyLoss(weight=weight)
    oracle.train()
    for i in range(STEPS):
        a, b, y = batch_of_step(i)
        opt.zero_grad()
        loss = crit(oracle(R.preprocess(a, "imagenet"), R.preprocess(b, "half")), y)
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
        if i % 10 == 0:
            print(i, losses[-1], flush=True)
    out = os.path.join(ROOT, "tests", "golden", out_name)
    d = {"steps": STEPS, "batch": B, "init_seed": SEED, "oracle": losses}
    d.update(extra)
    with open(out, "w") as f:
        json.dump(d, f)
    print("wrote", out)


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "learning"
    if mode == "random":
        run(lambda i: R.synthetic_batch(B, BATCH_SEED0 + i), torch.tensor([276 / 130, 276 / 146]),
            "loss_curve_200_seed1.json", {"batch_seed0": BATCH_SEED0})
        return
    rgb, th, y = R.synthetic_batch(SET_SIZE, SET_SEED)
    nb = SET_SIZE // B

    def batch(i):
        k = (i % nb) * B
        return rgb[k:k + B], th[k:k + B], y[k:k + B]

    run(batch, R.class_weights(y), "loss_curve_learning_200_seed1.json",
        {"set_size": SET_SIZE, "set_seed": SET_SEED, "data": "synthetic_batch(set_size, set_seed) cycled as %d batches of %d" % (nb, B)})


if __name__ == "__main__":
    main()
