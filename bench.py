#!/usr/bin/env python
"""Benchmark of the late-fusion DFU train step (BASELINE.json metric: fusion train pairs/s, % of bf16 tensor peak,
CPU baseline beside it).

  python bench.py --gpus 1 --steps K --warmup W            # config 3: B=128 pairs, bf16, one B200
  torchrun --nproc-per-node N ... bench.py --gpus N ...     # config 4: global batch 1024 pairs, NCCL all-reduce
  python bench.py --impl reference ...                      # the CPU fp32 PyTorch step (oracle) on the host cores

A "step" = zero_grad + forward + weighted CE + backward + AdamW on one synthetic batch (uint8 224x224x3 pairs,
random-init weights).  `value` is timed with the inputs already in HBM (CUDA-graph replay, CUDA events, max over
ranks); `e2e` goes through the public TrainStep API from pinned HOST buffers with the H2D copies and a D2H read of
the loss inside the timed region.  Prints ONE JSON line on rank 0."""
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

TRAIN_GFLOP_PER_PAIR = 129.443      # ResNet50 24.287 + ViT-B/16 105.147 + head 0.009 (fwd+bwd)
CONFIG3_BATCH = 128
CONFIG4_GLOBAL_BATCH = 1024


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_burst": d.get("bf16_tflops", 1590.0), "bf16_sustained": d.get("bf16_tflops_sustained", 1400.0),
                "hbm_gbs": d.get("hbm_gbs", 6650.0), "source": "MEASURED_PEAKS.json"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


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


def cpu_reference_run(steps, warmup, batch=16):
    """The reference algorithm (oracle port: torchvision ResNet50 + timm-layout ViT-B/16 + head, fp32, AdamW) on the
    host cores, all threads.  Bounded sample: `batch` pairs per step."""
    import torch
    from oracle import reference_model as R
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm runs on rank 0 alone and gets all usable cores back
    usable = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if "LOCAL_RANK" in os.environ and torch.get_num_threads() < usable:
        torch.set_num_threads(usable)
    threads = torch.get_num_threads()
    model = R.seeded_init_(R.OracleFusion(), 0, calibrate=False)
    rgb_u8, th_u8, labels = R.synthetic_batch(batch, 0)
    rgb, th = R.preprocess(rgb_u8, "imagenet"), R.preprocess(th_u8, "half")
    w = R.class_weights(labels)
    R.train_steps(model, [(rgb, th, labels)] * warmup, w)
    t0 = time.perf_counter()
    R.train_steps(model, [(rgb, th, labels)] * steps, w)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps * 1e3, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 5))
    warmup = max(1, min(args.warmup, 1))
    B = 16
    value, ms, threads = cpu_reference_run(steps, warmup, B)
    line = {
        "impl": "reference", "metric": "fusion_train_pairs_per_sec", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config 3: ResNet50+ViT-B/16 late-fusion train step, batch 128 pairs, 1 B200",
                   "global_batch": CONFIG3_BATCH, "per_gpu_batch": CONFIG3_BATCH, "image": "224x224x3 uint8 pairs",
                   "optimizer": "AdamW lr1e-4 wd1e-4", "loss": "class-weighted CE",
                   "sample": "each timed step = the same fp32 train step on %d pairs (bounded CPU sample of the workload)" % B},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": "oracle/reference_model.py fusion train step, B=%d, %d timed steps after %d warm-up, "
                                   "os.cpu_count()=%s" % (B, steps, warmup, os.cpu_count())},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


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


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--per-gpu-batch", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true")
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
        bucketer = GradBucketer(list(model.parameters()))
        opt.set_grad_source(bucketer.grad_source)
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
    gpu_launches = launches_per_step * args.steps if ts.graph is not None else ops.LAUNCHES - launches0

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
                    # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the ViT fc1 weight-gradient GEMM
                    # (3072x768x25216, CTA-pair kernel, fp32 split-K reduce-add) from the committed ncu --set full capture;
                    # algorithmic = dY 154.9 MB + X 38.7 MB + dW 9.4 MB
                    "traffic": {"bytes_per_launch": 207.9e6, "algorithmic_bytes_per_launch": 203.0e6,
                                "kernel": "dfu_tc_gemm_kernel<.,5,2>, ViT fc1 wgrad",
                                "source": "profiles/r01k_ncu_full_gemm_pair_fc1_wgrad.txt"},
                    "launches_per_step": n_tc, "attention_launches": n_attn, "tensor_ms_per_step": tc_ms,
                    "tensor_share_of_step": tc_ms / ms_per_step if ms_per_step else None,
                    "note": "achieved = %.3f GFLOP/pair x %d pairs / summed CUDA-event time (launching stream, GPU kept "
                            "busy) of the %d tcgen05 launches (GEMM, implicit-GEMM conv, attention) of one eager step; peak = "
                            "bf16_tflops_sustained of %s" % (TRAIN_GFLOP_PER_PAIR, B, n_tc, peaks["source"]),
                    "step_frac_of_peak_burst": value / n_gpus * TRAIN_GFLOP_PER_PAIR * 1e9 / (peaks["bf16_burst"] * 1e12),
                    "step_frac_of_peak_sustained": value / n_gpus * TRAIN_GFLOP_PER_PAIR * 1e9 / (peaks["bf16_sustained"] * 1e12)}
    if world > 1:
        dist.barrier()

    # ---------------------------------------------------------------- CPU baseline beside it (rank 0, N=1 only)
    cpu_baseline = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        v, ms, threads = cpu_reference_run(steps=3, warmup=1, batch=16)
        cpu_baseline = {"value": v, "unit": "pairs/s", "cores": threads, "kind": "port",
                        "sample": "oracle/reference_model.py fusion train step fp32, B=16, 3 timed steps after 1 warm-up "
                                  "(%.0f ms/step), os.cpu_count()=%s" % (ms, os.cpu_count())}

    if rank == 0:
        line = {
            "metric": "fusion_train_pairs_per_sec", "value": value, "unit": "pairs/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if n_gpus == 1 else "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": ("config 3: ResNet50+ViT-B/16 late-fusion train step, batch 128 pairs, 1 B200" if n_gpus == 1
                                    else "config 4: late-fusion data-parallel train step, global batch %d pairs, NCCL "
                                         "all-reduce of bf16 gradient buckets" % global_batch),
                       "global_batch": global_batch, "per_gpu_batch": B, "image": "224x224x3 uint8 pairs",
                       "optimizer": "AdamW lr1e-4 wd1e-4 (fused)", "loss": "class-weighted CE", "cuda_graph": ts.graph is not None,
                       "branch_overlap": not args.no_overlap, "parallelism": "dp%d" % n_gpus,
                       "l2": "per-step working set (~%.0f GB of activations) >> 126 MB L2; two input batches alternate"
                             % (torch.cuda.max_memory_allocated() / 2**30)},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "last_loss": loss_val},
            "gpu_launches": int(gpu_launches), "launches_per_step": int(launches_per_step),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
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
