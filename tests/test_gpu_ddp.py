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


def _channels_last(name, t):
    """DfuConv2d weights (and therefore their gradients / bucket slices) are stored (Cout, R, S, Cin); everything else as is."""
    return t.dim() == 4 and name.startswith("resnet.")


def _storage_flat(name, t):
    t = t.detach().float()
    return (t.permute(0, 2, 3, 1) if _channels_last(name, t) else t).contiguous().reshape(-1).cpu()


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
    named = list(model.named_parameters())
    torch.save({"params": {n: p.detach().cpu() for n, p in named},
                "buffers": {n: b.detach().cpu() for n, b in model.named_buffers()},
                # what this rank put on the wire (its own fp32 gradient, still in p.grad) and what came back (the bucket slice)
                "local_grads": {n: _storage_flat(n, p.grad) for n, p in named},
                "avg_grads": {n: bucketer.grad_source(p)[0].float().cpu().reshape(-1) for n, p in named},
                "loss": float(loss)}, os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    torch.cuda.synchronize()
    os._exit(0)        # NCCL communicator teardown can block behind captured/queued work; all ranks are in step here


@pytest.mark.parametrize("world", [2])
def test_n_gpu_step_equals_one_gpu_micro_batched_step(world, tmp_path):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs, %d visible" % (world, torch.cuda.device_count()))
    mp.spawn(_rank_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r)) for r in range(world)]
    from oracle import reference_model as R
    sd0 = R.seeded_init_(R.OracleFusion(), INIT_SEED, calibrate=False).state_dict()
    names = list(got[0]["params"])

    # (1) THE EXCHANGE, exactly: what every rank received == bf16( mean over ranks of bf16(its own gradient) ), within one bf16
    #     ulp of the fp32-summed value (NCCL averages in bf16), and identical on all ranks
    worst = 0.0
    for n in names:
        wire = [got[r]["local_grads"][n].to(torch.bfloat16).float() for r in range(world)]
        want = (sum(wire) / world).to(torch.bfloat16).float()
        for r in range(world):
            assert torch.equal(got[r]["avg_grads"][n], got[0]["avg_grads"][n]), "ranks received different averages: " + n
        lim = 2.0 ** -7 * max(float(want.abs().max()), 1e-12)
        err = float((got[0]["avg_grads"][n] - want).abs().max())
        worst = max(worst, err / lim)
        assert err <= lim, (n, err, lim)

    # (2) THE UPDATE, exactly: rank 0's parameters == one torch.optim.AdamW step from the initial weights on the averaged gradients
    dev = torch.device("cuda", 0)
    ref_params = [torch.nn.Parameter(sd0[n].clone().to(dev)) for n in names]
    ref_opt = torch.optim.AdamW(ref_params, lr=1e-4, weight_decay=1e-4)
    for p, n in zip(ref_params, names):
        g = got[0]["avg_grads"][n].to(dev)
        if _channels_last(n, p):               # bucket slices are in storage order: (Cout, R, S, Cin) for the ResNet convolutions
            p.grad = g.view(p.shape[0], p.shape[2], p.shape[3], p.shape[1]).permute(0, 3, 1, 2).contiguous()
        else:
            p.grad = g.view(p.shape)
    ref_opt.step()
    for p, n in zip(ref_params, names):
        assert torch.allclose(got[0]["params"][n].to(dev), p.detach(), rtol=1e-5, atol=2e-6), (
            n, (got[0]["params"][n].to(dev) - p.detach()).abs().max().item())
        assert torch.equal(got[0]["params"][n], got[1]["params"][n]), "replicas diverged at " + n

    # (3) AGAINST ONE GPU: the same `world` micro-batches one after the other on one device.  The runs are separate executions
    #     (fp32 atomics of the BatchNorm / split-K sums order differently, which bf16 rounding then amplifies), so this part is
    #     statistical: each rank saw its own micro-batch (loss), the averaged gradients point the same way.
    from dfu_multimodal_b200.functional import cross_entropy
    model = _build(dev)
    model.train()
    acc, losses, buffers0 = None, [], None
    for r in range(world):
        model.zero_grad(set_to_none=True)
        rgb, th, y = _batch(r, dev)
        loss = cross_entropy(model(rgb, th), y, None)
        loss.backward()
        torch.cuda.synchronize()
        losses.append(float(loss))
        g = {n: _storage_flat(n, p.grad).to(torch.bfloat16).float() for n, p in model.named_parameters()}
        acc = g if acc is None else {n: acc[n] + g[n] for n in acc}
        if r == 0:
            buffers0 = {n: b.detach().cpu().clone() for n, b in model.named_buffers()}
    for r in range(world):
        assert abs(got[r]["loss"] - losses[r]) < 2e-3 * max(1.0, abs(losses[r])), (r, got[r]["loss"], losses[r])
    cosines = {}
    for n in ["fusion.classifier.3.weight", "fusion.classifier.0.weight", "vit.blocks.11.mlp.fc2.weight", "vit.blocks.0.attn.qkv.weight",
              "resnet.layer4.2.conv3.weight", "resnet.layer1.0.conv1.weight", "resnet.conv1.weight"]:
        a, b = got[0]["avg_grads"][n], acc[n] / world
        cosines[n] = round(float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)), 4)
    print("DDP parity: exchange exact (worst %.2f of the one-ulp limit), update exact (<= 1e-6), replicas identical; "
          "gradient cosines against the 1-GPU micro-batched run: %s" % (worst, cosines))
    # (measured on 2 B200: 0.93-0.94 for the first ResNet convolutions at 4 pairs per rank, >= 0.986 elsewhere: two executions of the
    # SAME single-GPU step differ by as much)
    assert min(cosines.values()) > 0.9, cosines
    # BatchNorm statistics stay per rank: rank 0's are those of the 1-GPU model after micro-batch 0 alone
    # (two executions: the deep layers' batch statistics differ in the third digit, see the gradient cosines above)
    for n, b in buffers0.items():
        assert torch.allclose(got[0]["buffers"][n].float(), b.float(), rtol=2e-2, atol=2e-3), n
