"""Data parallelism on real GPUs (BASELINE config 4): an N-GPU step with per-GPU batch b
must equal the 1-GPU run that processes the same N micro-batches of b and averages their gradients -- BatchNorm statistics
stay per rank (the reference has no SyncBN), gradients travel as bf16 buckets and are averaged by NCCL.
Needs >= 2 visible GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp.py -m gpu`); skipped on one."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

B_PER_RANK = 4
INIT_SEED, DATA_SEED = 7, 900


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(dev):
    import dfu_multimodal_b200 as D
    from oracle import reference_model as R
    model = D.MultimodalFusionModel()
    model.load_state_dict(R.seeded_init_(R.OracleFusion(), INIT_SEED, calibrate=False).state_dict())
    for m in model.modules():
        if m.__class__.__name__ == "DfuDropout":
            m.p = 0.0
    return model.to(dev)


def _batch(rank, dev):
    from oracle import reference_model as R
    return [t.to(dev) for t in R.synthetic_batch(B_PER_RANK, DATA_SEED + rank)]


def _rank_main(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from dfu_multimodal_b200.ddp import GradBucketer, broadcast_parameters
    from dfu_multimodal_b200.optim import FusedAdamW
    from dfu_multimodal_b200.train import TrainStep
    model = _build(dev)
    broadcast_parameters(model)
    opt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
    bucketer = GradBucketer(list(model.parameters()))
    opt.set_grad_source(bucketer.grad_source, bucketer)
    ts = TrainStep(model, opt, None, use_graph=False, bucketer=bucketer)
    loss, _ = ts(*_batch(rank, dev))
    torch.cuda.synchronize()
    grads = {n: bucketer.grad_source(p)[0].float().cpu() for n, p in model.named_parameters()}
    torch.save({"params": {n: p.detach().cpu() for n, p in model.named_parameters()},
                "buffers": {n: b.detach().cpu() for n, b in model.named_buffers()},
                "avg_grads_flat": {n: g.reshape(-1) for n, g in grads.items()}, "loss": float(loss)},
               os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)        # NCCL communicator teardown can block behind captured/queued work; all ranks are in step here


@pytest.mark.parametrize("world", [2])
def test_n_gpu_step_equals_one_gpu_micro_batched_step(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, torch.cuda.device_count()))
    mp.spawn(_rank_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r)) for r in range(world)]

    # ---- the 1-GPU reference: the same `world` micro-batches one after the other, gradients cast to bf16 exactly as
    # the bucketer puts them on the wire, averaged, then ONE fused AdamW step
    from dfu_multimodal_b200.ddp import GradBucketer
    from dfu_multimodal_b200.functional import cross_entropy
    from dfu_multimodal_b200.optim import FusedAdamW
    dev = torch.device("cuda", 0)
    model = _build(dev)
    opt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
    bucketer = GradBucketer(list(model.parameters()))        # world size 1: casts into the bf16 buckets, no collective
    params = dict(model.named_parameters())
    acc, losses, buffers0 = None, [], None
    model.train()
    for r in range(world):
        model.zero_grad(set_to_none=True)
        bucketer.begin_step()
        rgb, th, y = _batch(r, dev)
        loss = cross_entropy(model(rgb, th), y, None)
        loss.backward()
        bucketer.finish_step()
        torch.cuda.synchronize()
        losses.append(float(loss))
        wire = {n: bucketer.grad_source(p)[0].float().clone() for n, p in params.items()}
        acc = wire if acc is None else {n: acc[n] + wire[n] for n in acc}
        if r == 0:
            buffers0 = {n: b.detach().cpu().clone() for n, b in model.named_buffers()}
    avg = {n: (g / world).to(torch.bfloat16) for n, g in acc.items()}      # fp32 sum of bf16 wire values, one rounding
    for n, p in params.items():
        bi, off, cnt = bucketer.slot[p]
        bucketer.flat[bi][off:off + cnt].copy_(avg[n])
    opt.set_grad_source(bucketer.grad_source, bucketer)
    opt.step()
    torch.cuda.synchronize()

    # every rank saw its own micro-batch
    for r in range(world):
        assert abs(got[r]["loss"] - losses[r]) < 1e-5, (r, got[r]["loss"], losses[r])
    # the averaged gradients handed to the optimizer: NCCL's bf16 AVG against the fp32-summed reference, within one bf16 ulp
    worst = 0.0
    for n in params:
        a, b = got[0]["avg_grads_flat"][n], avg[n].float().cpu().reshape(-1)
        lim = 2.0 ** -7 * max(float(b.abs().max()), 1e-12)
        worst = max(worst, float((a - b).abs().max()) / lim)
        assert float((a - b).abs().max()) <= lim, n
    # replicas stay identical, and equal the micro-batched run: AdamW's first update is lr * g/(|g| + eps), so a gradient
    # that differs in its last bf16 bit moves a weight by far less than lr = 1e-4
    n_elem, n_off = 0, 0
    for n, p in params.items():
        a0, a1, ref = got[0]["params"][n], got[1]["params"][n], p.detach().cpu()
        assert torch.equal(a0, a1), "replicas diverged at " + n
        d = (a0 - ref).abs()
        assert float(d.max()) <= 2.2e-4, (n, float(d.max()))
        n_elem += d.numel()
        n_off += int((d > 1e-6).sum())
    print("DDP parity: %d of %d elements differ by > 1e-6 (%.4f %%), worst gradient mismatch %.2f ulp-limits"
          % (n_off, n_elem, 100.0 * n_off / n_elem, worst))
    assert n_off / n_elem < 0.01
    # BatchNorm statistics stay per rank: rank 0's equal the 1-GPU model after micro-batch 0 alone
    for n, b in buffers0.items():
        assert torch.allclose(got[0]["buffers"][n].float(), b.float(), rtol=1e-5, atol=1e-6), n
