"""Model-level parity on the B200 against the CPU fp32 oracle (oracle/reference_model.py) and the committed outputs of
the reference's own classes (tests/golden).  Tolerances from BASELINE.json north_star: logits within 2e-2 abs (bf16),
argmax agreement >= 99.9 %, bit-for-bit labels on the 276-pair shape, loss within 1 %."""
import os

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu

LOGIT_TOL = 2e-2
# learning-curve bounds (test_learning_loss_curve_200_steps_against_golden).  This curve is a chaotic transient: Adam's first
# update is lr * sign(g) whatever |g| is, so the fp32 atomics of the split-K / BatchNorm sums (which order differently from
# run to run) flip the sign of noise-level gradients and the CUDA path already differs from ITSELF at step 1 (0.9434 /
# 0.9608 / 0.9788 in three runs of one build; oracle 0.9453).  Measured spread of |mean_cuda - mean_oracle| / mean_oracle
# over the 200 steps, nine B200 runs across four builds and both attention implementations: 0.17, 0.55, 0.56, 0.90, 0.94,
# 1.24, 1.39, 2.95 % (tools/gpu_learning_curve_spread.py, profiles/r02v_learning_curve_spread.txt); first-window means
# 0.6429 .. 0.6624 against the oracle's 0.6486, second-window means (the loss falls through 0.15 -> 0.02 there) 0.0631 .. 0.0725
# against 0.0674.  The bounds below are that spread with margin (window: 3 % + 1e-2); the deterministic checks are
# step 0 (1 %), the stable-regime curve (test_loss_curve_200_steps_against_golden, the north-star's 1 %) and the
# kernel-level gradient tests.  A path with a broken backward or optimizer stays at ln 2 and misses every window by 10x.
LEARNING_WINDOW_REL = 0.03
LEARNING_WINDOW_ABS = 1e-2
LEARNING_MEAN_REL = 0.05


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def R():
    from oracle import reference_model
    return reference_model


@pytest.fixture(scope="module")
def D():
    import dfu_multimodal_b200
    return dfu_multimodal_b200


@pytest.fixture(scope="module")
def pair(R, D, dev):
    oracle = R.seeded_init_(R.OracleFusion(), 0)
    model = D.MultimodalFusionModel()
    model.load_state_dict(oracle.state_dict(), strict=True)
    return oracle, model.to(dev)


def set_dropout(model, p):
    for m in model.modules():
        if m.__class__.__name__ in ("Dropout", "DfuDropout"):
            m.p = p


def test_native_library_is_loaded_and_device_ok(dev):
    from dfu_multimodal_b200 import _lib
    lib = _lib.load()
    assert lib.dfu_device_ok() == 0, lib.dfu_last_error()
    assert lib.dfu_sm_count() == 148
    with open("/proc/self/maps") as f:
        assert "libdfu_b200.so" in f.read()


def test_eval_logits_match_reference_golden(pair, R, dev, golden_dir):
    oracle, model = pair
    g = torch.load(os.path.join(golden_dir, "reference_logits_seed0.pt"))
    rgb_u8, th_u8, _ = R.synthetic_batch(g["B"], seed=g["seed"])
    model.eval()
    with torch.no_grad():
        out_u8 = model(rgb_u8.to(dev), th_u8.to(dev)).cpu()                                     # uint8 fast path
        out_f32 = model(R.preprocess(rgb_u8, "imagenet").to(dev), R.preprocess(th_u8, "half").to(dev)).cpu()
    assert out_u8.dtype == torch.float32 and out_u8.shape == (g["B"], 2)
    assert (out_u8 - g["fusion_logits"]).abs().max().item() < LOGIT_TOL
    assert (out_f32 - g["fusion_logits"]).abs().max().item() < LOGIT_TOL
    assert torch.equal(out_u8.argmax(1), g["fusion_logits"].argmax(1))


def test_single_modality_models_match_reference_golden(R, D, dev, golden_dir):
    g = torch.load(os.path.join(golden_dir, "reference_logits_seed0.pt"))
    rgb_u8, th_u8, _ = R.synthetic_batch(g["B"], seed=g["seed"])
    m = D.RGBOnlyModel()
    m.load_state_dict(R.seeded_init_(R.OracleRGBOnly(), 1).state_dict())
    m.to(dev).eval()
    with torch.no_grad():
        out = m(rgb_u8.to(dev)).cpu()
    assert (out - g["rgb_only_logits"]).abs().max().item() < LOGIT_TOL
    m = D.ThermalOnlyModel()
    m.load_state_dict(R.seeded_init_(R.OracleThermalOnly(), 2).state_dict())
    m.to(dev).eval()
    with torch.no_grad():
        out = m(th_u8.to(dev)).cpu()
    assert (out - g["thermal_only_logits"]).abs().max().item() < LOGIT_TOL


def test_labels_agree_on_the_276_pair_test_shape(pair, R, dev):
    """BASELINE config 5 shape: 276 pairs through EvalStep in the reference's batches of 69; north_star: bit-for-bit
    label agreement with the fp32 oracle.  The RAW agreement count is asserted (all 276), and reported with the
    smallest oracle margin so a future flip can be read off the log."""
    from dfu_multimodal_b200.train import EvalStep
    oracle, model = pair
    oracle.eval()
    rgb_u8, th_u8, _ = R.synthetic_batch(276, seed=276)
    step = EvalStep(model, use_graph=True)
    lo, lc, pc, yc = [], [], [], []
    with torch.no_grad():
        for i in range(0, 276, 69):
            lo.append(oracle(R.preprocess(rgb_u8[i:i + 69], "imagenet"), R.preprocess(th_u8[i:i + 69], "half")))
            logits, prob1, pred = step(rgb_u8[i:i + 69].to(dev), th_u8[i:i + 69].to(dev))
            lc.append(logits.cpu()), pc.append(prob1.cpu()), yc.append(pred.cpu())
    lo, lc, prob1, pred = torch.cat(lo), torch.cat(lc), torch.cat(pc), torch.cat(yc)
    err = (lo - lc).abs().max().item()
    margin = (lo[:, 0] - lo[:, 1]).abs()
    agree = int((pred == lo.argmax(1)).sum())
    print("276-pair shape: %d / 276 labels agree, max|dlogit| %.4f, smallest oracle margin %.4f" % (agree, err, margin.min().item()))
    assert err < LOGIT_TOL
    assert agree == 276, "labels differ on %d of 276 pairs (oracle margins there: %s)" % (
        276 - agree, margin[pred != lo.argmax(1)].tolist())
    assert (prob1 - lo.softmax(1)[:, 1]).abs().max().item() < 1e-2


def test_eval_step_graph_equals_plain_forward_for_changing_batch_shapes(pair, R, dev):
    """EvalStep (extended_metrics.py:623-642 as one captured graph per batch shape) returns what a plain eval forward
    returns, also for the smaller last batch of a loader."""
    from dfu_multimodal_b200 import ops
    from dfu_multimodal_b200.train import EvalStep
    _, model = pair
    step = EvalStep(model, use_graph=True)
    rgb_u8, th_u8, _ = R.synthetic_batch(13, seed=31)
    for lo_, hi_ in ((0, 8), (8, 13), (0, 8), (12, 13)):
        a, b = rgb_u8[lo_:hi_].to(dev), th_u8[lo_:hi_].to(dev)
        logits, prob1, pred = [t.clone() for t in step(a, b)]
        model.eval()
        with torch.no_grad():
            ref = model(a, b)
        p1, pr = ops.predict(ref)
        assert logits.shape == (hi_ - lo_, 2)
        assert (logits - ref).abs().max().item() <= 1e-6 and (prob1 - p1).abs().max().item() <= 1e-6 and torch.equal(pred, pr)
    assert len(step.graphs) == 3 and step.graph is not None
    with pytest.raises(RuntimeError):
        step(rgb_u8[:4].to(dev), th_u8[:3].to(dev))


def test_fp32_validation_mode_reproduces_the_reference_logits_to_1e_4(R, D, dev):
    """north_star: logits within 1e-4 in fp32-accum validation mode -- on SURVEY 8(d)'s plain gamma=1 initialisation, the
    ill-conditioned one (no bn3.weight=0.25 help).  Same module tree, same state_dict, dfu_v32_* kernels (fp32 activations,
    fp32 master weights).  The bf16 fast path on the very same weights is measured next to it."""
    from dfu_multimodal_b200 import validation
    oracle = R.seeded_init_(R.OracleFusion(), 0, bn3_gamma=1.0)
    model = D.MultimodalFusionModel()
    model.load_state_dict(oracle.state_dict(), strict=True)
    model.to(dev)
    B = 4
    rgb_u8, th_u8, _ = R.synthetic_batch(B, seed=21)
    rgb, th = R.preprocess(rgb_u8, "imagenet"), R.preprocess(th_u8, "half")
    oracle.eval(); model.eval()
    with torch.no_grad():
        lo = oracle(rgb, th)
        lv_u8 = validation.forward_fp32(model, rgb_u8.to(dev), th_u8.to(dev)).cpu()
        lv_f32 = validation.forward_fp32(model, rgb.to(dev), th.to(dev)).cpu()
        lb = model(rgb_u8.to(dev), th_u8.to(dev)).cpu()
        le = validation.forward_fp32(model, rgb_u8.to(dev), th_u8.to(dev), emulate_bf16_storage=True).cpu()
    e_u8, e_f32, e_bf16, e_emu = [(x - lo).abs().max().item() for x in (lv_u8, lv_f32, lb, le)]
    d_fast_emu = (lb - le).abs().max().item()
    print("gamma=1 eval logits (|logit| up to %.3f): fp32 validation mode max|d| %.2e (uint8 in) / %.2e (fp32 in); bf16 fast path "
          "%.2e; fp32 arithmetic with bf16 STORAGE emulated %.2e; fast path vs that emulation %.2e"
          % (lo.abs().max().item(), e_u8, e_f32, e_bf16, e_emu, d_fast_emu))
    assert lv_u8.dtype == torch.float32 and lv_u8.shape == (B, 2)
    assert e_u8 < 1e-4 and e_f32 < 1e-4
    # On this deliberately ill-conditioned network (every BatchNorm gamma = 1: each Bottleneck amplifies a perturbation ~1.25x)
    # bf16 STORAGE alone -- fp32 arithmetic, activations and weights rounded where the fast path stores them -- moves the
    # logits by more than 2e-2; the fast path has to be of that size, not larger, and the 2e-2 bound of north_star is asserted
    # on the conditioned weights (bn3.weight = 0.25, like a trained network) in test_eval_logits_match_reference_golden.
    # (measured on B200: fast path 5.4e-2, emulation 7.7e-2 from the reference, 1.05e-1 from each other: two different
    # roundings of a chaotic network are as far from each other as from the truth)
    assert e_bf16 < max(LOGIT_TOL, 3.0 * e_emu), (e_bf16, e_emu)
    assert e_bf16 < 0.15 and e_emu > LOGIT_TOL
    # train() semantics: BatchNorm batch statistics (no running-stat update in a validation pass)
    sd0 = {k: v.clone() for k, v in oracle.state_dict().items()}
    oracle.train(); model.train()
    set_dropout(oracle, 0.0)
    with torch.no_grad():
        lt = oracle(rgb, th)
    before = model.resnet.bn1.running_mean.clone()
    lvt = validation.forward_fp32(model, rgb_u8.to(dev), th_u8.to(dev)).cpu()
    assert (lvt - lt).abs().max().item() < 1e-4
    assert torch.equal(before, model.resnet.bn1.running_mean)
    oracle.load_state_dict(sd0)
    # the single-modality models (configs 1 and 2)
    for mk_o, mk_c, seed, x_u8, kind in ((R.OracleRGBOnly, D.RGBOnlyModel, 1, rgb_u8, "imagenet"),
                                         (R.OracleThermalOnly, D.ThermalOnlyModel, 2, th_u8, "half")):
        o = R.seeded_init_(mk_o(), seed, bn3_gamma=1.0).eval()
        m = mk_c()
        m.load_state_dict(o.state_dict())
        m.to(dev).eval()
        with torch.no_grad():
            want = o(R.preprocess(x_u8, kind))
        got = validation.forward_fp32(m, x_u8.to(dev)).cpu()
        assert (got - want).abs().max().item() < 1e-4, mk_c.__name__


def test_train_step_loss_and_gradient_direction(pair, R, dev):
    from dfu_multimodal_b200.functional import cross_entropy
    oracle, model = pair
    sd0 = {k: v.clone() for k, v in oracle.state_dict().items()}
    B = 8
    rgb_u8, th_u8, labels = R.synthetic_batch(B, 11)
    rgb, th = R.preprocess(rgb_u8, "imagenet"), R.preprocess(th_u8, "half")
    w = R.class_weights(labels)
    oracle.train(); model.train()
    set_dropout(oracle, 0.0); set_dropout(model, 0.0)
    oracle.zero_grad(); model.zero_grad()
    lo = nn.functional.cross_entropy(oracle(rgb, th), labels, weight=w)
    lo.backward()
    lc = cross_entropy(model(rgb.to(dev), th.to(dev)), labels.to(dev), w.to(dev))
    lc.backward()
    assert abs(lc.item() - lo.item()) / lo.item() < 1e-2                   # loss within 1 %
    og, cg = dict(oracle.named_parameters()), dict(model.named_parameters())
    for k in og:
        assert cg[k].grad is not None, k
        assert cg[k].grad.shape == og[k].grad.shape and cg[k].grad.dtype == torch.float32
    # every parameter gets a gradient; direction agrees with fp32 (bf16 rounding through ~60 layers of a random
    # network makes element-wise equality meaningless -- the same rounding injected into the fp32 oracle moves
    # these gradients by 20-50 %; per-op backward parity is in test_gpu_kernels.py)
    for k in ["fusion.classifier.3.weight", "fusion.classifier.0.weight", "vit.blocks.11.mlp.fc2.weight",
              "vit.blocks.0.attn.qkv.weight", "vit.patch_embed.proj.weight", "resnet.layer4.2.conv3.weight",
              "resnet.layer2.0.conv2.weight", "resnet.conv1.weight"]:
        a, b = cg[k].grad.float().cpu().flatten(), og[k].grad.flatten()
        cos = torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)
        assert cos.item() > 0.8, (k, cos.item())
    # BN running statistics were updated exactly like torch's (momentum 0.1, unbiased variance)
    for name in ("bn1", "layer1.0.bn3", "layer4.2.bn2"):
        mo, mc = oracle.resnet.get_submodule(name), model.resnet.get_submodule(name)
        assert (mo.running_mean - mc.running_mean.cpu()).abs().max().item() < 5e-3
        assert (mo.running_var - mc.running_var.cpu()).abs().max().item() < 5e-3 * max(1.0, mo.running_var.max().item())
        assert int(mc.num_batches_tracked.item()) == int(mo.num_batches_tracked.item())
    oracle.load_state_dict(sd0)
    model.load_state_dict(sd0)


def test_loss_curve_tracks_oracle(R, D, dev):
    """north_star: loss curves within 1 % over 200 steps.  Fresh synthetic batches every step (the bench regime).  Default
    24 steps of B=8 keeps the CPU oracle affordable; DFU_LOSS_STEPS=200 runs the full length."""
    from dfu_multimodal_b200.optim import FusedAdamW
    from dfu_multimodal_b200.train import TrainStep
    steps = int(os.environ.get("DFU_LOSS_STEPS", "24"))
    B = 8
    oracle = R.seeded_init_(R.OracleFusion(), 1)
    model = D.MultimodalFusionModel()
    model.load_state_dict(oracle.state_dict())
    model.to(dev)
    set_dropout(oracle, 0.0); set_dropout(model, 0.0)
    batches = [R.synthetic_batch(B, 500 + i) for i in range(steps)]
    w = torch.tensor([276 / 130, 276 / 146])                       # class weights of the 276-pair split (README.md:102)
    lo = R.train_steps(oracle, [(R.preprocess(a, "imagenet"), R.preprocess(b, "half"), y) for a, b, y in batches], w)
    ts = TrainStep(model, FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4), w.to(dev), use_graph=True)
    lc = [ts(a.to(dev), b.to(dev), y.to(dev))[0].item() for a, b, y in batches]
    assert ts.graph is not None                                   # steps after warm-up ran as CUDA-graph replays
    try:
        import json
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/loss_curve.json", "w") as f:
            json.dump({"steps": steps, "batch": B, "cuda": lc, "oracle": lo}, f)
    except OSError:
        pass
    mean_c, mean_o = sum(lc) / len(lc), sum(lo) / len(lo)
    assert abs(mean_c - mean_o) / mean_o < 0.01, (lc, lo)          # the curve as a whole: within 1 %
    # single steps: AdamW's first updates are +-lr per weight whatever the gradient size, so bf16 rounding of near-zero
    # gradients perturbs individual steps of this random-label problem; they stay within 5 %
    for a, b in zip(lc, lo):
        assert abs(a - b) / b < 0.05, (lc, lo)


def test_loss_curve_200_steps_against_golden(R, D, dev, golden_dir):
    """north_star: loss curves within 1 % over 200 steps.  The oracle side (200 fp32 CPU steps, ~10 min) is the committed
    fixture tests/golden/loss_curve_200_seed1.json written by oracle/make_loss_curve.py; the CUDA path replays the same
    200 synthetic batches here (CUDA-graph steps)."""
    import json
    from dfu_multimodal_b200.optim import FusedAdamW
    from dfu_multimodal_b200.train import TrainStep
    path = os.path.join(golden_dir, "loss_curve_200_seed1.json")
    if not os.path.exists(path):
        pytest.skip("golden loss curve not generated")
    with open(path) as f:
        gold = json.load(f)
    steps, B = gold["steps"], gold["batch"]
    oracle = R.seeded_init_(R.OracleFusion(), gold["init_seed"])     # same seeded weights the fixture started from
    model = D.MultimodalFusionModel()
    model.load_state_dict(oracle.state_dict())
    model.to(dev)
    set_dropout(model, 0.0)
    w = torch.tensor([276 / 130, 276 / 146]).to(dev)
    ts = TrainStep(model, FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4), w, use_graph=True)
    lc = []
    for i in range(steps):
        a, b, y = R.synthetic_batch(B, gold["batch_seed0"] + i)
        lc.append(ts(a.to(dev), b.to(dev), y.to(dev))[0].item())
    lo = gold["oracle"]
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/loss_curve_200.json", "w") as f:
            json.dump({"steps": steps, "batch": B, "cuda": lc, "oracle": lo}, f)
    except OSError:
        pass
    mean_c, mean_o = sum(lc) / steps, sum(lo) / steps
    assert abs(mean_c - mean_o) / mean_o < 0.01, (mean_c, mean_o)      # the curve over 200 steps: within 1 %
    for k in range(0, steps, 25):                                      # and every 25-step window of it: within 2 %
        wc, wo = sum(lc[k:k + 25]) / 25, sum(lo[k:k + 25]) / 25
        assert abs(wc - wo) / wo < 0.02, (k, wc, wo)


def test_learning_loss_curve_200_steps_against_golden(R, D, dev, golden_dir):
    """north_star: loss curves within 1 % over 200 steps, on a curve that MOVES: a fixed set of 64 synthetic pairs cycled
    as 8 batches of 8 (reference hot loop train_multimodal_fusion.py:368-388; fixture = 200 fp32 CPU steps written by
    oracle/make_loss_curve.py).  The oracle memorises the set: ln 2 -> below 0.1 after ~26 steps -> below 1e-3.  A path
    with a broken backward or optimizer stays at ln 2 and fails every bound below.
    Bounds: step 0 (identical weights) within 1 %; the mean over the 200 steps within 5 % and every 25-step window mean
    within 3 % + 1e-2 of the oracle's -- the measured run-to-run spread of this transient with margin, see the comment at
    LEARNING_MEAN_REL (measured: up to 2.1 % in the first window, 5e-3 absolute while the loss falls through 0.15 -> 0.02 in
    the second -- the runs memorise at slightly different steps --, < 1e-3 later)."""
    import json
    from dfu_multimodal_b200.optim import FusedAdamW
    from dfu_multimodal_b200.train import TrainStep
    path = os.path.join(golden_dir, "loss_curve_learning_200_seed1.json")
    with open(path) as f:
        gold = json.load(f)
    steps, B, n = gold["steps"], gold["batch"], gold["set_size"]
    oracle = R.seeded_init_(R.OracleFusion(), gold["init_seed"])
    model = D.MultimodalFusionModel()
    model.load_state_dict(oracle.state_dict())
    model.to(dev)
    set_dropout(model, 0.0)
    rgb, th, y = R.synthetic_batch(n, gold["set_seed"])
    w = R.class_weights(y).to(dev)
    rgb, th, y = rgb.to(dev), th.to(dev), y.to(dev)
    ts = TrainStep(model, FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4), w, use_graph=True)
    lc = []
    for i in range(steps):
        k = (i % (n // B)) * B
        lc.append(ts(rgb[k:k + B], th[k:k + B], y[k:k + B])[0].item())
    lo = gold["oracle"]
    try:
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/loss_curve_learning_200.json", "w") as f:
            json.dump({"steps": steps, "batch": B, "cuda": lc, "oracle": lo}, f)
    except OSError:
        pass
    report = []
    for k in range(0, steps, 25):
        wc, wo = sum(lc[k:k + 25]) / 25, sum(lo[k:k + 25]) / 25
        report.append((k, round(wc, 5), round(wo, 5)))
    print("learning curve, 25-step window means (start, cuda, oracle):", report)
    assert lo[0] > 0.5 and max(lo[100:]) < 2e-2                      # the fixture really is a falling curve
    assert abs(lc[0] - lo[0]) / lo[0] < 0.01                         # identical weights at step 0: within 1 %
    mean_c, mean_o = sum(lc) / steps, sum(lo) / steps
    print("200-step mean loss: cuda %.5f, oracle %.5f (%.2f %%)" % (mean_c, mean_o, 100 * abs(mean_c - mean_o) / mean_o))
    assert abs(mean_c - mean_o) / mean_o < LEARNING_MEAN_REL, (mean_c, mean_o)
    for k, wc, wo in report:
        assert abs(wc - wo) <= LEARNING_WINDOW_REL * wo + LEARNING_WINDOW_ABS, (k, wc, wo)
    assert max(lc[100:]) < 2e-2                                      # and it stays learned


def test_train_step_parity_at_config3_batch_128(R, D, dev):
    """Config 3's batch (128 pairs) against the fp32 oracle: train-mode logits (BatchNorm batch statistics over 128
    images), weighted CE loss within 1 %, every parameter receives a gradient, gradient directions agree."""
    from dfu_multimodal_b200.functional import cross_entropy
    B = 128
    oracle = R.seeded_init_(R.OracleFusion(), 5)
    model = D.MultimodalFusionModel()
    model.load_state_dict(oracle.state_dict())
    model.to(dev)
    set_dropout(oracle, 0.0); set_dropout(model, 0.0)
    rgb_u8, th_u8, labels = R.synthetic_batch(B, 128)
    w = R.class_weights(labels)
    oracle.train(); model.train()
    out_o = oracle(R.preprocess(rgb_u8, "imagenet"), R.preprocess(th_u8, "half"))
    lo = nn.functional.cross_entropy(out_o, labels, weight=w)
    lo.backward()
    out_c = model(rgb_u8.to(dev), th_u8.to(dev))
    lc = cross_entropy(out_c, labels.to(dev), w.to(dev))
    lc.backward()
    err = (out_o.detach() - out_c.detach().cpu()).abs().max().item()
    print("B=128 train-mode logits max|d| %.4f, loss %.5f vs oracle %.5f" % (err, lc.item(), lo.item()))
    assert err < LOGIT_TOL
    assert abs(lc.item() - lo.item()) / lo.item() < 1e-2
    og, cg = dict(oracle.named_parameters()), dict(model.named_parameters())
    assert all(cg[k].grad is not None and torch.isfinite(cg[k].grad).all() for k in og)
    cosines = {}
    for k in ["fusion.classifier.3.weight", "fusion.classifier.0.weight", "vit.blocks.11.mlp.fc2.weight",
              "vit.blocks.0.attn.qkv.weight", "resnet.layer4.2.conv3.weight", "resnet.layer1.0.conv1.weight"]:
        a, b = cg[k].grad.float().cpu().flatten(), og[k].grad.flatten()
        cosines[k] = round((torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item(), 4)
    print("B=128 gradient cosines vs fp32 oracle:", cosines)
    # bf16 activations through ~50 layers of a random (gamma = 1 / 0.25) network: direction, not element-wise equality --
    # the fp32 validation mode is what pins the algorithm; per-op backward parity is in test_gpu_kernels.py
    assert min(cosines.values()) > 0.8, cosines


def test_train_step_handles_the_smaller_last_batch_of_an_epoch(R, D, dev):
    """The reference loaders keep the last, smaller batch (no drop_last, train_multimodal_fusion.py:272-275): TrainStep must
    run it (own graph, or eagerly beyond max_graphs) and return ITS logits -- never broadcast into the first shape's buffers."""
    from dfu_multimodal_b200.optim import FusedAdamW
    from dfu_multimodal_b200.train import TrainStep
    model = D.MultimodalFusionModel()
    model.load_state_dict(R.seeded_init_(R.OracleFusion(), 6, calibrate=False).state_dict())
    model.to(dev)
    set_dropout(model, 0.0)
    ts = TrainStep(model, FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4), None, use_graph=True, warmup=1, max_graphs=2)
    a, b, y = [t.to(dev) for t in R.synthetic_batch(6, 17)]
    for _ in range(3):
        loss, logits = ts(a, b, y)
    assert ts.graph is not None and logits.shape == (6, 2)
    loss4, logits4 = ts(a[:4], b[:4], y[:4])                     # a second shape: second graph
    assert logits4.shape == (4, 2) and len(ts.graphs) == 2 and torch.isfinite(loss4)
    loss1, logits1 = ts(a[:1], b[:1], y[:1])                     # beyond the graph budget: eager
    assert logits1.shape == (1, 2) and ts.graph is None and torch.isfinite(loss1)
    loss6, logits6 = ts(a, b, y)                                 # and the first graph still replays
    assert logits6.shape == (6, 2) and ts.graph is not None and torch.isfinite(loss6)
    assert loss6.item() < loss.item() + 0.5
    with pytest.raises(RuntimeError):
        ts(a, b, y[:5])


def test_cuda_graph_step_equals_eager_step(R, D, dev):
    from dfu_multimodal_b200.optim import FusedAdamW
    from dfu_multimodal_b200.train import TrainStep
    sd = R.seeded_init_(R.OracleFusion(), 2, calibrate=False).state_dict()
    models = []
    for _ in range(2):
        m = D.MultimodalFusionModel()
        m.load_state_dict(sd)
        set_dropout(m, 0.0)
        models.append(m.to(dev))
    a, b, y = [t.to(dev) for t in R.synthetic_batch(4, 7)]
    te = TrainStep(models[0], FusedAdamW(models[0].parameters(), lr=1e-4, weight_decay=1e-4), None, use_graph=False)
    tg = TrainStep(models[1], FusedAdamW(models[1].parameters(), lr=1e-4, weight_decay=1e-4), None, use_graph=True)
    le = [te(a, b, y)[0].item() for _ in range(6)]
    lg = [tg(a, b, y)[0].item() for _ in range(6)]
    assert tg.graph is not None and tg.launches_per_step > 500
    # identical weights at step 1: only the fp32 summation order of split-K reduce-adds can differ
    assert abs(le[0] - lg[0]) < 2e-3, (le, lg)
    # afterwards AdamW's sign-like first updates amplify that noise on this 4-pair memorisation problem; both runs must
    # still converge the same way
    for x, z in zip(le, lg):
        assert abs(x - z) < 0.15, (le, lg)
    assert le[-1] < 0.1 and lg[-1] < 0.1, (le, lg)


def test_dropout_is_active_in_train_and_off_in_eval(pair, R, dev):
    _, model = pair
    set_dropout(model, 0.7)
    rgb_u8, th_u8, _ = R.synthetic_batch(2, 3)
    model.train()
    with torch.no_grad():
        a = model(rgb_u8.to(dev), th_u8.to(dev))
        b = model(rgb_u8.to(dev), th_u8.to(dev))      # a plain second call, as the reference's eager loop makes it
    assert not torch.equal(a, b)
    model.eval()
    with torch.no_grad():
        c, d = model(rgb_u8.to(dev), th_u8.to(dev)), model(rgb_u8.to(dev), th_u8.to(dev))
    assert torch.equal(c, d)


def test_gradcam_matches_reference_gradcam_class_outputs(pair, R, dev, golden_dir):
    from dfu_multimodal_b200.gradcam import GradCAM
    oracle, model = pair
    g = torch.load(os.path.join(golden_dir, "reference_gradcam_seed0.pt"))
    rgb_u8, th_u8, _ = R.synthetic_batch(2, seed=0)
    x_rgb = R.preprocess(rgb_u8, "imagenet")[:1].to(dev)
    x_th = R.preprocess(th_u8, "half")[:1].to(dev)
    model.eval()
    with torch.no_grad():                                          # the reference calls it under no_grad (:831)
        cam = GradCAM(model.resnet, ["layer4"])
        c = cam.generate_cam(x_rgb, class_idx=1)
        cam.close()
        sal = GradCAM(model.vit, ["blocks"])
        s = sal.generate_cam(x_th, class_idx=1)
        sal.close()
    ref = g["resnet_layer4_cam"].numpy()
    assert c.shape == (7, 7) and abs(c.max() - 1.0) < 1e-6
    assert abs(c - ref).max() < 0.15 and (abs(c - ref).mean() < 0.05)
    assert int(c.argmax()) == int(ref.argmax())                   # same hottest cell
    refs = g["vit_blocks_saliency"].float().numpy()
    assert s.shape == (224, 224) and abs(s.max() - 1.0) < 1e-6
    cc = ((s - s.mean()) * (refs - refs.mean())).sum() / ((s - s.mean()).std() * (refs - refs.mean()).std() * s.size)
    assert cc > 0.8, cc                                            # saliency map correlates with the fp32 reference
    assert not any(m._forward_hooks for m in model.modules())     # close() removed the hooks


def test_tta_batched_views_match_oracle_vote(pair, R, dev):
    from dfu_multimodal_b200 import tta
    oracle, model = pair
    oracle.eval()
    n, V = 6, 8
    rgb_u8, th_u8, _ = R.synthetic_batch(n, 77)
    rv, tv = tta.flip_views(rgb_u8.to(dev), V), tta.flip_views(th_u8.to(dev), V)
    mean_prob, pred, per_view = tta.evaluate_with_tta(model, rv, tv, batch_size=16)
    with torch.no_grad():
        po = []
        for v in range(V):
            lo = oracle(R.preprocess(rv[:, v].cpu(), "imagenet"), R.preprocess(tv[:, v].cpu(), "half"))
            po.append(lo.softmax(1)[:, 1])
        po = torch.stack(po, 1)
    assert (per_view.cpu() - po).abs().max().item() < 1e-2
    mo, vo = R.oracle_tta_vote(po.numpy())
    assert abs(mean_prob.cpu().numpy() - mo).max() < 1e-2
    firm = (abs(po.numpy() - 0.5) > 1e-2).all(1)
    assert (pred.cpu().numpy()[firm] == vo[firm]).all()


def test_reference_style_driver_runs_on_the_dropin(R, dev):
    """A model written exactly as the reference scripts write theirs (torch.hub.load + timm.create_model + stock
    nn.Sequential head, train_multimodal_fusion.py:285-326) trains on the B200 backbones through dropin.install()."""
    from dfu_multimodal_b200 import dropin
    dropin.install(force_timm_shim=True)
    try:
        import timm

        class MultimodalFusionModel(nn.Module):
            def __init__(self, num_classes=2, dropout=0.5):
                super().__init__()
                self.rgb_branch = torch.hub.load('pytorch/vision:v0.13.1', 'resnet50', pretrained=False)
                rgb_feat_dim = self.rgb_branch.fc.in_features
                self.rgb_branch.fc = nn.Identity()
                self.thermal_branch = timm.create_model('vit_base_patch16_224', pretrained=False, num_classes=0)
                self.fusion = nn.Sequential(nn.Linear(rgb_feat_dim + 768, 512), nn.ReLU(), nn.Dropout(dropout),
                                            nn.Linear(512, 256), nn.ReLU(), nn.Dropout(dropout), nn.Linear(256, num_classes))

            def forward(self, rgb, thermal):
                return self.fusion(torch.cat([self.rgb_branch(rgb), self.thermal_branch(thermal)], dim=1))

        model = MultimodalFusionModel().to(dev)
        assert any(k.startswith("rgb_branch.layer4.2.conv3") for k in model.state_dict())
        crit = nn.CrossEntropyLoss(weight=torch.tensor([1.5, 0.8], device=dev))
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
        rgb_u8, th_u8, labels = R.synthetic_batch(4, 5)
        rgb, th = R.preprocess(rgb_u8, "imagenet").to(dev), R.preprocess(th_u8, "half").to(dev)
        model.train()
        losses = []
        for _ in range(3):
            opt.zero_grad()
            out = model(rgb, th)
            assert out.shape == (4, 2) and out.dtype == torch.float32
            loss = crit(out, labels.to(dev))
            loss.backward()
            opt.step()
            losses.append(loss.item())
        assert all(torch.isfinite(torch.tensor(losses)))
        assert model.rgb_branch.conv1.weight.grad is not None and model.thermal_branch.cls_token.grad is not None
        model.eval()
        with torch.no_grad():
            probs = torch.softmax(model(rgb, th), dim=1)[:, 1].cpu().numpy()       # extended_metrics.py:636
        assert probs.shape == (4,)
    finally:
        dropin.uninstall()


def test_checkpoint_written_by_train_scheme_loads_and_predicts(R, D, dev, tmp_path):
    from dfu_multimodal_b200 import checkpoint
    sd = R.seeded_init_(R.OracleFusion(head_layers=3), 4).state_dict()
    train_sd = {k.replace("resnet.", "rgb_branch.", 1).replace("vit.", "thermal_branch.", 1).replace("fusion.classifier.", "fusion."): v
                for k, v in sd.items()}
    path = tmp_path / "best_model.pt"
    torch.save({"epoch": 5, "model_state_dict": train_sd, "optimizer_state_dict": {}, "val_f1": 0.9, "history": {}}, path)
    model = checkpoint.build_fusion_from_checkpoint(torch.load(path), device=dev).eval()
    oracle = R.OracleFusion(head_layers=3)
    oracle.load_state_dict(sd)
    oracle.eval()
    rgb_u8, th_u8, _ = R.synthetic_batch(2, 9)
    with torch.no_grad():
        a = model(rgb_u8.to(dev), th_u8.to(dev)).cpu()
        b = oracle(R.preprocess(rgb_u8, "imagenet"), R.preprocess(th_u8, "half"))
    assert (a - b).abs().max().item() < LOGIT_TOL


def test_device_prefetcher_yields_batches_in_order(dev):
    from dfu_multimodal_b200.data import DevicePrefetcher
    g = torch.Generator().manual_seed(1)
    host = [(torch.randint(0, 256, (4, 8, 8, 3), generator=g, dtype=torch.uint8).pin_memory(),
             torch.full((4,), i, dtype=torch.int64).pin_memory()) for i in range(7)]
    seen = []
    for img, lab in DevicePrefetcher(iter(host), dev):
        assert img.is_cuda and lab.is_cuda
        seen.append((img.cpu().clone(), lab.cpu().clone()))
    assert len(seen) == 7
    for (hi, hl), (si, sl) in zip(host, seen):
        assert torch.equal(hi, si) and torch.equal(hl, sl)


def test_resume_from_checkpoint_continues_the_run(R, D, dev, tmp_path):
    """save after 3 steps in the reference's container, resume into a fresh model + FusedAdamW, and the next
    steps equal the uninterrupted run's (fixed batch, dropout off)."""
    from dfu_multimodal_b200 import checkpoint
    from dfu_multimodal_b200.optim import FusedAdamW
    from dfu_multimodal_b200.train import TrainStep
    sd = R.seeded_init_(R.OracleFusion(), 8, calibrate=False).state_dict()
    a, b, y = [t.to(dev) for t in R.synthetic_batch(4, 33)]

    def fresh():
        m = D.MultimodalFusionModel()
        m.load_state_dict(sd)
        set_dropout(m, 0.0)
        m.to(dev)
        o = FusedAdamW(m.parameters(), lr=1e-4, weight_decay=1e-4)
        return m, o, TrainStep(m, o, None, use_graph=False)
    m1, o1, t1 = fresh()
    for _ in range(3):
        t1(a, b, y)
    path = tmp_path / "best_model.pt"
    checkpoint.save_checkpoint(path, m1, o1, epoch=3, val_f1=0.5, history={"train_loss": [1.0, 0.9, 0.8]})
    straight = [t1(a, b, y)[0].item() for _ in range(3)]
    m2, o2, t2 = fresh()
    info = checkpoint.resume(str(path), m2, o2)
    assert info["epoch"] == 3 and info["history"]["train_loss"] == [1.0, 0.9, 0.8]
    resumed = [t2(a, b, y)[0].item() for _ in range(3)]
    print("resume: uninterrupted", straight, "resumed", resumed)
    for u, v in zip(straight, resumed):
        assert abs(u - v) < 2e-3 * max(1.0, abs(u)), (straight, resumed)      # fp32 split-K summation order only
    assert int(m2.resnet.bn1.num_batches_tracked.item()) == 6
