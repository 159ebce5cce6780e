"""Test-time augmentation (a12).  The reference runs N x num_tta batch-1 forwards with a host sync each
(notebooks/test_time_augmentation.py:331-401); here all views of all samples go through the model as large batches
and the mean-probability / majority-vote reduction is one kernel.  Vote rule as the reference: per view
pred_v = prob_v > 0.5, final pred = mean(pred_v) > 0.5 (ties -> 0), prob = mean(prob_v)."""
import torch

from . import ops


@torch.no_grad()
def evaluate_with_tta(model, rgb_views, thermal_views, batch_size=256):
    """rgb_views / thermal_views: (N, V, ...) device tensors (uint8 NHWC or fp32 NCHW per view).
    Returns (mean_prob [N] fp32, pred [N] int64, per_view_prob [N,V])."""
    model.eval()
    N, V = rgb_views.shape[:2]
    rgb = rgb_views.reshape(N * V, *rgb_views.shape[2:])
    thermal = thermal_views.reshape(N * V, *thermal_views.shape[2:])
    probs = torch.empty(N * V, dtype=torch.float32, device=rgb.device)
    for i in range(0, N * V, batch_size):
        logits = model(rgb[i:i + batch_size].contiguous(), thermal[i:i + batch_size].contiguous())
        p1, _ = ops.predict(logits)          # softmax[:,1] for 2 logits, sigmoid for the 1-logit TTA head
        probs[i:i + batch_size] = p1
    mean_prob, pred = ops.tta_vote(probs.view(N, V))
    return mean_prob, pred, probs.view(N, V)


def flip_views(img_u8, num_views):
    """Deterministic light views of uint8 NHWC images (identity, h-flip, v-flip, both, cycled): a stand-in for the
    reference's random PIL augmentation, which is CPU-side and out of scope."""
    views = []
    for v in range(num_views):
        x = img_u8
        if v & 1:
            x = x.flip(2)
        if v & 2:
            x = x.flip(1)
        views.append(x)
    return torch.stack(views, 1).contiguous()
