#!/usr/bin/env python
"""Benchmark of the late-fusion DFU classifier on B200 (BASELINE.json metric: fusion train pairs/s, % of bf16 tensor
peak, CPU baseline beside it).  Prints ONE JSON line on rank 0.

  python bench.py --gpus 1 --steps K --warmup W            # config 3 (the headline): B=128 pairs, bf16, one B200
  torchrun --nproc-per-node N ... bench.py --gpus N ...     # config 4: global batch 1024 pairs, NCCL all-reduce
  python bench.py --config 1|2|5                            # ResNet50 B=32 inference | ViT B=64 inference + train |
                                                            # 276 pairs x (forward + Grad-CAM on both branches + 8 TTA views)
  python bench.py --impl reference [--config c]             # the reference algorithm (oracle port, fp32) on the HOST cores
  python bench.py --impl torch_cuda [--config c]            # stock PyTorch (cuDNN / cuBLAS / SDPA, bf16 autocast,
                                                            # channels_last, fused AdamW) on the same B200: SURVEY.md 8(d)
                                                            # "reference on this box" -- the reference ships no kernels

A config-3 "step" = zero_grad + forward + weighted CE + backward + AdamW on one synthetic batch (uint8 224x224x3 pairs,
random-init weights).  `value` is timed with the inputs already in HBM (CUDA-graph replay, CUDA events, max over ranks);
`e2e` goes through the public TrainStep API from pinned HOST buffers with the H2D copies and a D2H read of the loss inside
the timed region.  The default config-3 line also carries `stock_pytorch_cuda` (the same step in stock PyTorch on the same
GPU, measured in the same process right after ours)."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY.md 8(d) / BASELINE.md section 2: algorithmic work
TRAIN_GFLOP_PER_PAIR = 129.443      # ResNet50 24.287 + ViT-B/16 105.147 + head 0.009 (fwd+bwd)
INFER_GFLOP_PER_PAIR = 43.303
RESNET_FWD_GFLOP, RESNET_TRAIN_GFLOP = 8.174, 24.287
VIT_FWD_GFLOP, VIT_TRAIN_GFLOP = 35.126, 105.147
CONFIG1_BATCH, CONFIG2_BATCH, CONFIG3_BATCH, CONFIG4_GLOBAL_BATCH, CONFIG5_PAIRS, CONFIG5_VIEWS = 32, 64, 128, 1024, 276, 8


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_burst": d.get("bf16_tflops", 1590.0), "bf16_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "MEASURED_PEAKS.json"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def roofline_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    summary (profiles/roofline_traffic.json is written by tools/ncu_summary.py traffic <report>); None when absent."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples = []
        self.reasons = set()
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(float(parts[0]))
                    self.max_mhz = float(parts[1])
                    for n, v in zip(names, parts[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(0.2)

    def result(self):
        self.stop_flag.set()
        self.join(timeout=6)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


_JSON_FD = None


def _claim_stdout():
    """Keep stdout for the ONE JSON line: file descriptor 1 is pointed at stderr for the rest of the process, so banners
    that libraries write to stdout (NCCL's version line under NCCL_DEBUG=VERSION, for one) cannot precede it."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


WORKLOADS = {
    1: "config 1: RGB-only ResNet50 2-class inference, batch 32, 224x224 synthetic, random init",
    2: "config 2: thermal-only ViT-Base/16 2-class inference + train step, batch 64, 1 B200",
    3: "config 3: ResNet50+ViT-B/16 late-fusion train step, batch 128 pairs, 1 B200",
    5: "config 5: fusion inference + Grad-CAM (both branches) + test-time augmentation (8 views) over 276 synthetic pairs",
}
METRICS = {1: ("rgb_inference_images_per_sec", "images/s"), 2: ("vit_train_images_per_sec", "images/s"),
           3: ("fusion_train_pairs_per_sec", "pairs/s"), 5: ("fusion_eval_gradcam_tta_pairs_per_sec", "pairs/s")}


# ======================================================================================== CPU arm (oracle port, fp32)
def _cpu_threads():
    import torch
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone and gets all usable cores back
    usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if "LOCAL_RANK" in os.environ and torch.get_num_threads() < usable:
        torch.set_num_threads(usable)
    return torch.get_num_threads()


def cpu_reference_run(config, steps, warmup):
    """The reference algorithm (oracle port: torchvision ResNet50 + timm-layout ViT-B/16 + head, fp32, AdamW) on the host
    cores, all threads, on a BOUNDED sample of the config's workload.  Returns (value, ms_per_step, threads, sample)."""
    import torch
    from oracle import reference_model as R
    threads = _cpu_threads()
    if config == 3:
        B = 16
        model = R.seeded_init_(R.OracleFusion(), 0, calibrate=False)
        rgb_u8, th_u8, labels = R.synthetic_batch(B, 0)
        rgb, th = R.preprocess(rgb_u8, "imagenet"), R.preprocess(th_u8, "half")
        w = R.class_weights(labels)
        R.train_steps(model, [(rgb, th, labels)] * warmup, w)
        t0 = time.perf_counter()
        R.train_steps(model, [(rgb, th, labels)] * steps, w)
        dt = time.perf_counter() - t0
        return B * steps / dt, dt / steps * 1e3, threads, "fusion train step fp32 on %d pairs per step (config 3 runs 128)" % B
    if config == 1:
        B = CONFIG1_BATCH
        model = R.seeded_init_(R.OracleRGBOnly(), 1).eval()
        x = R.preprocess(R.synthetic_batch(B, 0)[0], "imagenet")
        with torch.no_grad():
            for _ in range(warmup):
                model(x)
            t0 = time.perf_counter()
            for _ in range(steps):
                model(x).softmax(1)
            dt = time.perf_counter() - t0
        return B * steps / dt, dt / steps * 1e3, threads, "ResNet50 inference fp32, the full batch of %d" % B
    if config == 2:
        B = 8
        model = R.seeded_init_(R.OracleThermalOnly(), 2)
        _, th_u8, labels = R.synthetic_batch(B, 0)
        th = R.preprocess(th_u8, "half")
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
        model.train()

        def step():
            opt.zero_grad()
            torch.nn.functional.cross_entropy(model(th), labels).backward()
            opt.step()
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
        return B * steps / dt, dt / steps * 1e3, threads, "ViT-B/16 train step fp32 on %d images per step (config 2 runs 64)" % B
    if config == 5:
        n, V = 4, CONFIG5_VIEWS
        model = R.seeded_init_(R.OracleFusion(), 0).eval()
        rgb_u8, th_u8, _ = R.synthetic_batch(n, 276)

        def job():
            rgb, th = R.preprocess(rgb_u8, "imagenet"), R.preprocess(th_u8, "half")
            with torch.no_grad():
                model(rgb, th).softmax(1)
            for i in range(n):                         # reference: per-sample Grad-CAM on both branches
                R.oracle_gradcam(model.resnet, ["layer4"], rgb[i:i + 1])
                R.oracle_gradcam(model.vit, ["blocks"], th[i:i + 1])
            with torch.no_grad():
                for v in range(V):
                    model(rgb.flip(3) if v & 1 else rgb, th.flip(3) if v & 1 else th).softmax(1)
        for _ in range(min(warmup, 1)):
            job()
        steps = max(1, min(steps, 2))
        t0 = time.perf_counter()
        for _ in range(steps):
            job()
        dt = time.perf_counter() - t0
        return n * steps / dt, dt / steps * 1e3, threads, ("forward + Grad-CAM on both branches + %d TTA views, fp32, on %d "
                                                           "pairs per pass (config 5 runs 276)" % (V, n))
    raise SystemExit("no CPU arm for config %s" % config)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warmup = max(1, min(args.warmup, 1))
    cfg = args.config
    value, ms, threads, sample = cpu_reference_run(cfg, steps, warmup)
    metric, unit = METRICS[cfg]
    line = {
        "impl": "reference", "metric": metric, "value": value, "unit": unit,
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[cfg], "image": "224x224x3 uint8", "optimizer": "AdamW lr1e-4 wd1e-4",
                   "loss": "class-weighted CE", "sample": "each timed step = " + sample + " (bounded CPU sample)"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port",
                         "sample": "oracle/reference_model.py, %s, %d timed steps after %d warm-up, os.cpu_count()=%s"
                                   % (sample, steps, warmup, os.cpu_count())},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ======================================================================================== stock PyTorch on the same GPU
def _stock_models(R, torch):
    """The oracle modules (torchvision ResNet50 + timm-layout ViT) with the attention of the ViT routed through
    F.scaled_dot_product_attention, as timm >= 0.9 does on CUDA -- stock PyTorch at its best."""
    import torch.nn.functional as F

    def sdpa_forward(self, x):
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, C // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v, scale=self.scale)
        return self.proj(x.transpose(1, 2).reshape(B, N, C))
    if not hasattr(R._Attention, "_oracle_forward"):
        R._Attention._oracle_forward = R._Attention.forward
    R._Attention.forward = sdpa_forward


def _restore_oracle(R):
    if hasattr(R._Attention, "_oracle_forward"):
        R._Attention.forward = R._Attention._oracle_forward


def stock_pytorch_cuda_train(dev, B, steps, warmup, which="fusion"):
    """One B200, stock PyTorch: bf16 autocast, channels_last ResNet, SDPA attention, cudnn.benchmark, fused AdamW.
    Returns a dict with the device-resident value and the reference-style end-to-end loop (H2D + loss.item() per step)."""
    import torch
    from oracle import reference_model as R
    _stock_models(R, torch)
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    if which == "fusion":
        model = R.seeded_init_(R.OracleFusion(dropout=0.7), 0, calibrate=False)
    else:
        model = R.seeded_init_(R.OracleThermalOnly(), 2, calibrate=False)
    model = model.to(dev).to(memory_format=torch.channels_last).train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4, fused=True)
    rgb_u8, th_u8, labels = R.synthetic_batch(B, 0)
    cw = R.class_weights(labels).to(dev)
    host = [R.preprocess(rgb_u8, "imagenet").contiguous(memory_format=torch.channels_last).pin_memory(),
            R.preprocess(th_u8, "half").pin_memory(), labels.pin_memory()]
    if which != "fusion":
        host = host[1:]
    devb = [t.to(dev) for t in host]

    def step(*batch):
        *x, y = batch
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = model(*x)
        loss = torch.nn.functional.cross_entropy(out.float(), y, weight=cw if which == "fusion" else None)
        loss.backward()
        opt.step()
        return loss

    for _ in range(max(warmup, 3)):
        step(*devb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step(*devb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    # end to end, written like the reference hot loop (train_multimodal_fusion.py:368-388): .to(DEVICE) per batch, .item() per step
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step(*[t.to(dev, non_blocking=True) for t in host])
        loss.item()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    mem = torch.cuda.max_memory_allocated() / 2**30
    del model, opt
    torch.cuda.empty_cache()
    _restore_oracle(R)
    return {"value": B / ms * 1e3, "unit": "pairs/s" if which == "fusion" else "images/s", "ms_per_step": ms,
            "e2e": {"value": B * steps / e2e_s, "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host),
                    "d2h_bytes_per_step": 4},
            "steps": steps, "batch": B, "peak_mem_gib": mem,
            "what": "oracle module (torchvision ResNet50 + timm-layout ViT-B/16) in stock PyTorch %s on the same GPU: bf16 "
                    "autocast, channels_last, F.scaled_dot_product_attention, cudnn.benchmark, torch.optim.AdamW(fused=True), "
                    "fp32 normalised inputs" % torch.__version__}


def stock_pytorch_cuda_infer(dev, which, B, steps, warmup):
    import torch
    from oracle import reference_model as R
    _stock_models(R, torch)
    torch.backends.cudnn.benchmark = True
    mk = {"rgb": lambda: R.seeded_init_(R.OracleRGBOnly(), 1), "thermal": lambda: R.seeded_init_(R.OracleThermalOnly(), 2),
          "fusion": lambda: R.seeded_init_(R.OracleFusion(), 0)}[which]
    model = mk().to(dev).to(memory_format=torch.channels_last).eval()
    rgb_u8, th_u8, _ = R.synthetic_batch(B, 0)
    rgb = R.preprocess(rgb_u8, "imagenet").to(dev).contiguous(memory_format=torch.channels_last)
    th = R.preprocess(th_u8, "half").to(dev)
    x = {"rgb": (rgb,), "thermal": (th,), "fusion": (rgb, th)}[which]
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(max(warmup, 3)):
            model(*x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            model(*x).float().softmax(1)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model
    torch.cuda.empty_cache()
    _restore_oracle(R)
    return {"value": B / ms * 1e3, "ms_per_step": ms, "batch": B, "steps": steps,
            "what": "stock PyTorch %s eval forward, bf16 autocast, channels_last, SDPA, cudnn.benchmark" % torch.__version__}


def run_torch_cuda(args):
    """`--impl torch_cuda`: the comparator alone (rank 0 only under torchrun)."""
    import torch
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = args.config
    sampler = ClockSampler(0)
    sampler.start()
    extra = {}
    if cfg == 3:
        r = stock_pytorch_cuda_train(dev, CONFIG3_BATCH, args.steps, args.warmup, "fusion")
        flops = TRAIN_GFLOP_PER_PAIR
    elif cfg == 2:
        r = stock_pytorch_cuda_train(dev, CONFIG2_BATCH, args.steps, args.warmup, "thermal")
        extra["inference"] = stock_pytorch_cuda_infer(dev, "thermal", CONFIG2_BATCH, args.steps, args.warmup)
        flops = VIT_TRAIN_GFLOP
    elif cfg == 1:
        r = stock_pytorch_cuda_infer(dev, "rgb", CONFIG1_BATCH, args.steps, args.warmup)
        r["e2e"] = {"value": r["value"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        flops = RESNET_FWD_GFLOP
    else:
        raise SystemExit("--impl torch_cuda covers configs 1, 2, 3")
    clocks = sampler.result()
    peaks = measured_peaks()
    metric, unit = METRICS[cfg]
    line = {"impl": "torch_cuda", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": 1, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOADS[cfg], "what": r["what"]},
            "e2e": dict(r["e2e"], unit=unit), "gpu_launches": 0, "clocks": clocks,
            "tflops_per_gpu": r["value"] * flops / 1e3,
            "frac_of_bf16_burst_peak": r["value"] * flops * 1e9 / (peaks["bf16_burst"] * 1e12)}
    line.update(extra)
    _emit(line)


# ======================================================================================== our arm: configs 1, 2, 5
def _time_events(torch, fn, steps, warmup):
    for _ in range(max(warmup, 3)):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def run_small_config(args):
    """Configs 1, 2, 5 on one B200 through the package's public API (EvalStep / TrainStep / gradcam / tta)."""
    import torch
    import dfu_multimodal_b200 as D
    from dfu_multimodal_b200 import ops, tta
    from dfu_multimodal_b200.gradcam import GradCAM
    from dfu_multimodal_b200.optim import FusedAdamW
    from dfu_multimodal_b200.train import EvalStep, TrainStep
    from oracle import reference_model as R          # synthetic data + seeded weights + the CPU arm only

    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    cfg = args.config
    peaks = measured_peaks()
    metric, unit = METRICS[cfg]
    extra = {}
    sampler = ClockSampler(0)

    def host_pair(n, seed):
        rgb, th, y = R.synthetic_batch(n, seed)
        return rgb.pin_memory(), th.pin_memory(), y.pin_memory()

    if cfg == 1:
        B = CONFIG1_BATCH
        model = D.RGBOnlyModel()
        model.load_state_dict(R.seeded_init_(R.OracleRGBOnly(), 1).state_dict())
        model.to(dev)
        step = EvalStep(model)
        hosts = [host_pair(B, s)[0] for s in range(2)]
        devs = [h.to(dev) for h in hosts]
        it = [0]

        def fn():
            it[0] += 1
            return step(devs[it[0] % 2])
        step(devs[0]); step(devs[0])
        sampler.start()
        ms = _time_events(torch, fn, args.steps, args.warmup)
        clocks = sampler.result()
        launches = (step.launches_per_step or 0) * args.steps
        # end to end: pinned host uint8 -> H2D -> forward -> probabilities and labels back on the host
        out_host = torch.empty(B, dtype=torch.float32).pin_memory()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.steps):
            x = hosts[i % 2].to(dev, non_blocking=True)
            _, prob1, pred = step(x)
            out_host.copy_(prob1, non_blocking=True)
            torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        value, per_step, flops = B / ms * 1e3, B, RESNET_FWD_GFLOP
        e2e = {"value": B * args.steps / e2e_s, "unit": unit, "h2d_bytes_per_step": hosts[0].numel(), "d2h_bytes_per_step": 4 * B}
        if not args.no_torch_baseline:
            extra["stock_pytorch_cuda"] = stock_pytorch_cuda_infer(dev, "rgb", B, args.steps, args.warmup)
    elif cfg == 2:
        B = CONFIG2_BATCH
        model = D.ThermalOnlyModel()
        model.load_state_dict(R.seeded_init_(R.OracleThermalOnly(), 2).state_dict())
        model.to(dev)
        hosts = [host_pair(B, s)[1:] for s in range(2)]
        devs = [tuple(t.to(dev) for t in h) for h in hosts]
        ev = EvalStep(model)
        it = [0]

        def fn_inf():
            it[0] += 1
            return ev(devs[it[0] % 2][0])
        ms_inf = _time_events(torch, fn_inf, args.steps, args.warmup)
        ts = TrainStep(model, FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4), None, use_graph=True)
        for i in range(5):
            ts(*devs[i % 2])

        def fn_tr():
            it[0] += 1
            return ts(*devs[it[0] % 2])
        sampler.start()
        ms = _time_events(torch, fn_tr, args.steps, args.warmup)
        clocks = sampler.result()
        launches = (ts.launches_per_step or 0) * args.steps
        loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.steps):
            loss, _ = ts(*[t.to(dev, non_blocking=True) for t in hosts[i % 2]])
            loss_host.copy_(loss, non_blocking=True)
            torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        value, per_step, flops = B / ms * 1e3, B, VIT_TRAIN_GFLOP
        e2e = {"value": B * args.steps / e2e_s, "unit": unit, "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in hosts[0]),
               "d2h_bytes_per_step": 4}
        extra["inference"] = {"value": B / ms_inf * 1e3, "unit": "images/s", "ms_per_step": ms_inf,
                              "tflops": B / ms_inf * VIT_FWD_GFLOP, "frac_of_bf16_burst_peak": B / ms_inf * VIT_FWD_GFLOP / peaks["bf16_burst"]}
        if not args.no_torch_baseline:
            extra["stock_pytorch_cuda"] = stock_pytorch_cuda_train(dev, B, args.steps, args.warmup, "thermal")
            extra["stock_pytorch_cuda"]["inference"] = stock_pytorch_cuda_infer(dev, "thermal", B, args.steps, args.warmup)
    else:
        # config 5: reference loops extended_metrics.py:623-642, grad_cam_visualization.py:561-632,
        # test_time_augmentation.py:331-401 over the 276-pair test split
        N, V = CONFIG5_PAIRS, CONFIG5_VIEWS
        model = D.MultimodalFusionModel()
        model.load_state_dict(R.seeded_init_(R.OracleFusion(), 0).state_dict())
        model.to(dev).eval()
        rgb_h, th_h, _ = host_pair(N, 276)
        ev = EvalStep(model)
        cam_r, cam_t = GradCAM(model.resnet, ["layer4"]), GradCAM(model.vit, ["blocks"])
        phases = {"forward": 0.0, "gradcam": 0.0, "tta": 0.0}
        results = {}

        def job(timed):
            t = [time.perf_counter()]

            def lap(name):
                torch.cuda.synchronize()
                t.append(time.perf_counter())
                if timed:
                    phases[name] += t[-1] - t[-2]
            rgb, th = rgb_h.to(dev, non_blocking=True), th_h.to(dev, non_blocking=True)
            probs, preds = [], []
            for i in range(0, N, 69):
                _, p1, pr = ev(rgb[i:i + 69], th[i:i + 69])
                probs.append(p1.clone()), preds.append(pr.clone())
            results["prob"] = torch.cat(probs).cpu()
            lap("forward")
            cams, sals = [], []
            for i in range(0, N, 69):
                x_rgb = ops.nhwc_bf16_to_nchw_f32(ops.preprocess_u8_nhwc(rgb[i:i + 69], "imagenet"))
                x_th = ops.nhwc_bf16_to_nchw_f32(ops.preprocess_u8_nhwc(th[i:i + 69], "half"))
                cams += cam_r.generate_cam_batch(x_rgb)
                sals += cam_t.generate_cam_batch(x_th)
            results["cams"] = len(cams) + len(sals)
            lap("gradcam")
            rv, tv = tta.flip_views(rgb, V), tta.flip_views(th, V)
            mean_prob, pred, _ = tta.evaluate_with_tta(model, rv, tv, batch_size=276)
            results["tta"] = mean_prob.cpu()
            lap("tta")
        job(False)
        l0 = ops.LAUNCHES
        job(False)
        launches_job = ops.LAUNCHES - l0
        steps = max(1, min(args.steps, 5))
        sampler.start()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            job(True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        clocks = sampler.result()
        cam_r.close(), cam_t.close()
        args.steps = steps
        ms, value, per_step = dt * 1e3, N / dt, N
        launches = launches_job * steps
        # work: 1 clean forward + 8 TTA forwards of the fusion model, Grad-CAM = forward+backward of each branch
        flops = (1 + V) * INFER_GFLOP_PER_PAIR + RESNET_TRAIN_GFLOP + VIT_TRAIN_GFLOP
        e2e = {"value": value, "unit": unit, "h2d_bytes_per_step": rgb_h.numel() + th_h.numel(),
               "d2h_bytes_per_step": 4 * N * 2 + N * (49 + 224 * 224) * 4,
               "note": "the job IS end to end: pinned host uint8 images in, probabilities / heat maps / votes back on the host"}
        extra["phases_ms"] = {k: v / steps * 1e3 for k, v in phases.items()}
        extra["maps_per_job"] = results["cams"]

    cpu_baseline = None
    if not args.no_cpu_baseline:
        v, cms, threads, sample = cpu_reference_run(cfg, 2 if cfg != 1 else 3, 1)
        cpu_baseline = {"value": v, "unit": unit, "cores": threads, "kind": "port",
                        "sample": "oracle/reference_model.py, %s (%.0f ms per step), os.cpu_count()=%s" % (sample, cms, os.cpu_count())}
    tf = value * flops / 1e3
    line = {"metric": metric, "value": value, "unit": unit, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": {"workload": WORKLOADS[cfg], "units_per_step": per_step, "image": "224x224x3 uint8",
                                            "l2": "two input batches alternate" if cfg != 5 else "276-pair job >> L2"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": tf, "peak": peaks["bf16_burst"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_burst"],
                         "traffic": None, "note": "whole step: %.3f GFLOP per unit x units/s; peak = bf16_tflops (burst) of %s"
                                                  % (flops, peaks["source"])},
            "cpu_baseline": cpu_baseline}
    line.update(extra)
    _emit(line)


# ======================================================================================== our arm: configs 3 / 4
def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch_cuda"])
    ap.add_argument("--config", type=int, default=3, choices=[1, 2, 3, 5],
                    help="BASELINE.json configs (4 = config 3's step under torchrun, chosen by --gpus N)")
    ap.add_argument("--per-gpu-batch", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true", help="skip the stock-PyTorch-on-this-GPU comparator")
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--ddp-defer", action="store_true", help="N > 1: start the gradient all-reduces after backward instead of beside it")
    ap.add_argument("--bucket-mb", type=int, default=0, help="N > 1: gradient bucket size in MB (0 = ddp.BUCKET_BYTES)")
    ap.add_argument("--vit-sms", type=int, default=0, help="SM budget of the ViT branch stream (0 = whole device)")
    ap.add_argument("--resnet-sms", type=int, default=0, help="SM budget of the ResNet branch stream")
    ap.add_argument("--watchdog", type=int, default=0, help="dump all Python stacks to stderr after this many seconds")
    args = ap.parse_args()
    if args.watchdog > 0:
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog, exit=True)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA path has no CPU fallback (use --impl reference for the CPU arm)")
    if args.impl == "torch_cuda":
        run_torch_cuda(args)
        return
    if args.config != 3:
        run_small_config(args)
        return
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    import dfu_multimodal_b200 as D
    from dfu_multimodal_b200 import ops
    from dfu_multimodal_b200.ddp import GradBucketer, broadcast_parameters
    from dfu_multimodal_b200.optim import FusedAdamW
    from dfu_multimodal_b200.train import TrainStep
    from oracle import reference_model as R          # synthetic data generator + CPU baseline only

    if args.per_gpu_batch > 0:
        B = args.per_gpu_batch
    else:
        B = CONFIG3_BATCH if n_gpus == 1 else CONFIG4_GLOBAL_BATCH // n_gpus
    global_batch = B * n_gpus

    torch.manual_seed(0)
    model = D.MultimodalFusionModel(num_classes=2, dropout=0.7).to(dev)
    if not args.no_overlap:
        model.enable_branch_overlap(True, args.vit_sms, args.resnet_sms)
    broadcast_parameters(model)
    opt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
    bucketer = None
    if world > 1:
        bucketer = GradBucketer(list(model.parameters()), overlap=not args.ddp_defer,
                                **({"bucket_bytes": args.bucket_mb << 20} if args.bucket_mb else {}))
        opt.set_grad_source(bucketer.grad_source, bucketer)
    # synthetic data: 2 distinct batches alternate so no step re-reads the previous step's inputs from L2
    host = []
    for s in range(2):
        rgb, th, y = R.synthetic_batch(B, seed=1000 * rank + s)
        host.append((rgb.pin_memory(), th.pin_memory(), y.pin_memory()))
    devb = [tuple(t.to(dev) for t in hb) for hb in host]
    cw = R.class_weights(torch.cat([hb[2] for hb in host])).to(dev)
    ts = TrainStep(model, opt, class_weights=cw, use_graph=not args.no_graph, bucketer=bucketer)

    # graph capture / allocator warm-up (not part of the W warm-up steps)
    for i in range(5):
        ts(*devb[i % 2])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident timing ("value")
    for i in range(max(args.warmup, 3)):
        ts(*devb[i % 2])
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = ops.LAUNCHES
    e0.record()
    for i in range(args.steps):
        ts(*devb[i % 2])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.result() if sampler else None
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = global_batch / ms_per_step * 1e3
    launches_per_step = ts.launches_per_step or 0
    gpu_launches = ops.LAUNCHES - launches0           # graph replays count their captured kernels (train.TrainStep)

    # ---------------------------------------------------------------- end to end from host buffers ("e2e")
    barrier()
    e2e_steps = args.steps
    from dfu_multimodal_b200.data import DevicePrefetcher
    # every step's inputs come from pinned HOST memory (H2D inside the timed region, issued one step ahead on a copy
    # stream by the package's DevicePrefetcher) and every step's loss is read back to the host (D2H + sync)
    # ... through a pinned 4-byte buffer: the copy is enqueued right behind the step and its value is taken one step later,
    # so reading the result does not drain the launch queue (the reference's per-step .item(), :383, does)
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]

    prefetcher = DevicePrefetcher((), dev)       # one copy stream and one pair of device slots for the whole run

    def e2e_run(n_steps, stamps=None):
        loss_ready = [None, None]
        n_read, val = 0, 0.0
        prefetcher.batches = (host[j % 2] for j in range(n_steps))
        for i, db in enumerate(prefetcher):
            loss, _ = ts(*db)
            loss_host[i % 2].copy_(loss, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            loss_ready[i % 2] = ev
            if i > 0:
                loss_ready[(i - 1) % 2].synchronize()
                val = float(loss_host[(i - 1) % 2])
                n_read += 1
                if stamps is not None:
                    stamps.append(time.perf_counter())
        loss_ready[(n_steps - 1) % 2].synchronize()
        val = float(loss_host[(n_steps - 1) % 2])
        n_read += 1
        if stamps is not None:
            stamps.append(time.perf_counter())
        assert n_read == n_steps
        return val

    e2e_run(max(2, min(args.warmup, 3)))       # untimed: first use of the copy stream, its device slots and the pinned path
    torch.cuda.synchronize()
    barrier()
    stamps = []
    t0 = time.perf_counter()
    loss_val = e2e_run(e2e_steps, stamps)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = global_batch * e2e_steps / e2e_s
    if rank == 0 and len(stamps) > 1:
        gaps = sorted((b - a) * 1e3 for a, b in zip([t0] + stamps[:-1], stamps))
        sys.stderr.write("e2e per-step host gaps (ms): median %.2f  max %.2f  first %.2f\n" % (
            gaps[len(gaps) // 2], gaps[-1], (stamps[0] - t0) * 1e3))
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    d2h = 4

    # ---------------------------------------------------------------- roofline of the dominant kernel (tcgen05 GEMM)
    roofline = None
    if True:     # every rank runs the two eager steps (they contain the gradient all-reduces); rank 0 reports
        peaks = measured_peaks()
        # one eager step with CUDA events around every GEMM launch, on the launching stream
        model.enable_branch_overlap(False)       # serial streams, so each event pair times exactly one launch
        eager = TrainStep(model, opt, class_weights=cw, use_graph=False, bucketer=bucketer)
        eager(*devb[0])
        torch.cuda.synchronize()
        ops.GEMM_EVENTS = []
        # keep the GPU busy (~0.15 s spin) while the host enqueues the eager step, so that the event pairs bracket
        # kernels that run back to back and measure device time, not launch gaps
        torch.cuda._sleep(int(0.15 * 1.9e9))
        eager(*devb[1])
        torch.cuda.synchronize()
        evs, ops.GEMM_EVENTS = ops.GEMM_EVENTS, None
        tc_ms = sum(ev[0].elapsed_time(ev[1]) for ev in evs)
        n_tc = len(evs)
        n_attn = sum(1 for ev in evs if str(ev[2][-1]).startswith("attn"))
        flops_step = TRAIN_GFLOP_PER_PAIR * 1e9 * B
        achieved = flops_step / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else None
        roofline = {"bound": "tensor", "kernel": "dfu_tc_gemm_kernel (+ attn_fwd/attn_bwd): every tcgen05 launch of the step",
                    "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                    "frac": (achieved / peaks["bf16_sustained"]) if achieved else None,
                    # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, read from the
                    # committed ncu --set full summary (profiles/roofline_traffic.json)
                    "traffic": roofline_traffic(),
                    "launches_per_step": n_tc, "attention_launches": n_attn, "tensor_ms_per_step": tc_ms,
                    "tensor_share_of_step": tc_ms / ms_per_step if ms_per_step else None,
                    "note": "achieved = %.3f GFLOP/pair x %d pairs / summed CUDA-event time (launching stream, GPU kept "
                            "busy) of the %d tcgen05 launches (GEMM, implicit-GEMM conv, attention) of one eager step; peak = "
                            "bf16_tflops_sustained of %s" % (TRAIN_GFLOP_PER_PAIR, B, n_tc, peaks["source"]),
                    "step_frac_of_peak_burst": value / n_gpus * TRAIN_GFLOP_PER_PAIR * 1e9 / (peaks["bf16_burst"] * 1e12),
                    "step_frac_of_peak_sustained": value / n_gpus * TRAIN_GFLOP_PER_PAIR * 1e9 / (peaks["bf16_sustained"] * 1e12)}
    if world > 1:
        dist.barrier()

    # ---------------------------------------------------------------- stock PyTorch on the same GPU (rank 0, N=1 only)
    stock = None
    if rank == 0 and n_gpus == 1 and not args.no_torch_baseline:
        mem_ours = torch.cuda.max_memory_allocated() / 2**30
        try:
            stock = stock_pytorch_cuda_train(dev, B, min(args.steps, 10), 3, "fusion")
            stock["speedup_of_this_repo"] = {"device_resident": value / stock["value"], "e2e": e2e_value / stock["e2e"]["value"]}
        except Exception as exc:  # noqa: BLE001 -- the comparator must never take the headline measurement down
            stock = {"error": "%s: %s" % (type(exc).__name__, exc)}
    else:
        mem_ours = torch.cuda.max_memory_allocated() / 2**30

    # ---------------------------------------------------------------- CPU baseline beside it (rank 0, N=1 only)
    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        v, ms, threads, sample = cpu_reference_run(3, steps=3, warmup=1)
        cpu_baseline = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                        "sample": "oracle/reference_model.py %s, 3 timed steps after 1 warm-up (%.0f ms/step), "
                                  "os.cpu_count()=%s" % (sample, ms, os.cpu_count())}

    if rank == 0:
        line = {
            "metric": "fusion_train_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if n_gpus == 1 else "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": (WORKLOADS[3] if n_gpus == 1
                                    else "config 4: late-fusion data-parallel train step, global batch %d pairs, NCCL "
                                         "all-reduce of bf16 gradient buckets" % global_batch),
                       "global_batch": global_batch, "per_gpu_batch": B, "image": "224x224x3 uint8 pairs",
                       "optimizer": "AdamW lr1e-4 wd1e-4 (fused)", "loss": "class-weighted CE", "cuda_graph": ts.graph is not None,
                       "branch_overlap": not args.no_overlap, "parallelism": "dp%d" % n_gpus,
                       "l2": "per-step working set (~%.0f GB of activations) >> 126 MB L2; two input batches alternate" % mem_ours},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "last_loss": loss_val},
            "gpu_launches": int(gpu_launches), "launches_per_step": int(launches_per_step),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "stock_pytorch_cuda": stock,
            "tflops_per_gpu": value / n_gpus * TRAIN_GFLOP_PER_PAIR / 1e3,
        }
        _emit(line)
    if world > 1:
        # The captured CUDA graphs hold NCCL work on the communicator: tearing the process group down while they are alive
        # blocks (measured: destroy_process_group never returns).  All ranks are done and in step here, so leave directly.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
