"""Run-to-run and implementation-to-implementation spread of the 200-step learning curve (tests/test_gpu_model.py)."""
import json, os, sys
sys.path.insert(0, os.getcwd())
import torch
import dfu_multimodal_b200 as D
from dfu_multimodal_b200 import ops
from dfu_multimodal_b200.optim import FusedAdamW
from dfu_multimodal_b200.train import TrainStep
from oracle import reference_model as R

gold = json.load(open("tests/golden/loss_curve_learning_200_seed1.json"))
steps, B, n = gold["steps"], gold["batch"], gold["set_size"]
lo = gold["oracle"]
dev = torch.device("cuda")
for impl in sys.argv[1:]:
    ops.set_attention_impl(impl)
    oracle = R.seeded_init_(R.OracleFusion(), gold["init_seed"])
    model = D.MultimodalFusionModel()
    model.load_state_dict(oracle.state_dict())
    model.to(dev)
    for m in model.modules():
        if hasattr(m, "p") and m.__class__.__name__.lower().find("dropout") >= 0:
            m.p = 0.0
    rgb, th, y = R.synthetic_batch(n, gold["set_seed"])
    w = R.class_weights(y).to(dev)
    rgb, th, y = rgb.to(dev), th.to(dev), y.to(dev)
    ts = TrainStep(model, FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4), w, use_graph=True)
    lc = []
    for i in range(steps):
        k = (i % (n // B)) * B
        lc.append(ts(rgb[k:k + B], th[k:k + B], y[k:k + B])[0].item())
    wins = [(round(sum(lc[k:k + 25]) / 25, 5), round(sum(lo[k:k + 25]) / 25, 5)) for k in range(0, steps, 25)]
    print(impl, "first steps", [round(x, 4) for x in lc[:6]], "mean diff %.2f %%" % (100 * abs(sum(lc) - sum(lo)) / sum(lo)), "windows", wins, flush=True)
