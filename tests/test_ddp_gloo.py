"""Data-parallel host logic on CPU: world_size-2 gloo run of the gradient bucketer (bf16 buckets, hooks, all-reduce,
grad_source hand-off).  The cast kernel is injected as a torch op here because the product cast is CUDA-only."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from dfu_multimodal_b200.ddp import GradBucketer


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _cpu_cast(src, dst):
    dst.copy_(src.detach().reshape(-1).to(torch.bfloat16))


def _model():
    torch.manual_seed(0)
    return nn.Sequential(nn.Linear(16, 32), nn.ReLU(), nn.Linear(32, 32), nn.ReLU(), nn.Linear(32, 2))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = _model()
        bucketer = GradBucketer(list(model.parameters()), bucket_bytes=2048, cast_fn=_cpu_cast)
        assert len(bucketer.buckets) >= 2
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn(8, 16, generator=g)
        y = torch.randint(0, 2, (8,), generator=g)
        for step in range(2):                       # twice: begin_step must reset the bookkeeping
            model.zero_grad(set_to_none=True)
            bucketer.begin_step()
            nn.functional.cross_entropy(model(x), y).backward()
            bucketer.finish_step()
        res = {}
        for name, p in model.named_parameters():
            gsrc, is_bf16, scale = bucketer.grad_source(p)
            assert is_bf16 and gsrc.dtype == torch.bfloat16 and gsrc.numel() == p.numel()
            res[name] = (gsrc.float() * scale).view(p.shape).clone()
        if rank == 0:
            torch.save(res, out)
    finally:
        dist.destroy_process_group()


def test_bucketed_bf16_allreduce_matches_mean_of_rank_gradients(tmp_path):
    out = str(tmp_path / "grads.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    # reference: mean over the two ranks' gradients, computed in one process
    ref = None
    for rank in range(2):
        model = _model()
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn(8, 16, generator=g)
        y = torch.randint(0, 2, (8,), generator=g)
        nn.functional.cross_entropy(model(x), y).backward()
        grads = {n: p.grad.clone() for n, p in model.named_parameters()}
        ref = grads if ref is None else {n: ref[n] + grads[n] for n in ref}
    for n in ref:
        want = ref[n] / 2
        assert (got[n] - want).abs().max().item() <= 2e-2 * max(1e-3, want.abs().max().item()) + 1e-4, n


def test_bucket_layout_is_aligned_and_reverse_ordered():
    model = _model()
    b = GradBucketer(list(model.parameters()), bucket_bytes=1 << 20, cast_fn=_cpu_cast)
    params = list(model.parameters())
    assert b.buckets[0][0] is params[-1]            # last registered parameter first (backward order)
    for p in params:
        bi, off, n = b.slot[p]
        assert off % 8 == 0 and n == p.numel()      # 16-byte aligned bf16 slices
    b.remove()


def test_parameter_without_gradient_is_skipped_not_given_a_stale_one():
    """A parameter that gets no gradient in a step must not be handed last step's reduced gradient."""
    model = _model()
    extra = nn.Linear(32, 2)
    params = list(model.parameters()) + list(extra.parameters())
    b = GradBucketer(params, bucket_bytes=1 << 20, cast_fn=_cpu_cast)
    x = torch.randn(4, 16)
    h = model[:-1](x)
    b.begin_step()
    (model[-1](h) + extra(h)).sum().backward()          # step 1: everyone gets a gradient
    b.finish_step()
    assert all(b.grad_source(p) is not None for p in params)
    assert float(b.grad_source(extra.weight)[0].float().abs().max()) > 0
    for p in params:
        p.grad = None
    b.begin_step()
    model(x).sum().backward()                            # step 2: `extra` is not used
    b.finish_step()
    assert b.grad_source(extra.weight) is None and b.grad_source(extra.bias) is None
    bi, off, n = b.slot[extra.weight]
    assert float(b.flat[bi][off:off + n].float().abs().max()) == 0.0
    assert b.grad_source(model[0].weight) is not None
    b.remove()
