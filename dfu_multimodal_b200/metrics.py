"""Device-side metrics and asynchronous loss logging.

The reference reads the loss and the predictions back to the host after EVERY step (`loss.item()`, `predicted.cpu()`,
`labels.cpu()`, notebooks/train_multimodal_fusion.py:383-386; `probs.cpu()` per batch in extended_metrics.py:636-640) and
computes its metrics with sklearn at the end (MedicalMetricsCalculator, extended_metrics.py:374-446).  `DeviceMetrics`
keeps everything on the GPU: `update()` is one small kernel per batch with no synchronisation, `compute()` one kernel plus
ONE device->host copy per epoch, returning the reference's metric names."""
import torch

from . import _lib, ops
from ._lib import check

NAMES = ["tn", "fp", "fn", "tp", "accuracy", "precision", "recall", "f1", "sensitivity", "specificity", "ppv", "npv", "fpr", "fnr",
         "auc_roc", "auc_pr", "mcc", "kappa", "balanced_accuracy", "f_harmonic", "mean_loss", "n", "overflowed"]


def _p(t):
    return None if t is None else t.data_ptr()


class DeviceMetrics:
    def __init__(self, device, capacity=4096):
        if not (0 < capacity <= 16384):
            raise ValueError("capacity must be in 1..16384 samples per epoch (one CTA sorts the scores)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("DeviceMetrics lives on the GPU (no CPU path); got %s" % (device,))
        self.capacity = capacity
        self.state = torch.zeros(6, dtype=torch.int64, device=self.device)
        self.loss_acc = torch.zeros(1, dtype=torch.float64, device=self.device)
        self.probs = torch.empty(capacity, dtype=torch.float32, device=self.device)
        self.y_true = torch.empty(capacity, dtype=torch.int32, device=self.device)
        self.y_pred = torch.empty(capacity, dtype=torch.int32, device=self.device)
        self.out = torch.empty(24, dtype=torch.float64, device=self.device)

    def reset(self):
        """Start of an epoch: two dfu_zero launches, no sync."""
        L = _lib.load()
        st = torch.cuda.current_stream().cuda_stream
        check(L.dfu_zero(_p(self.state), self.state.numel() * 8, st), "dfu_zero")
        check(L.dfu_zero(_p(self.loss_acc), 8, st), "dfu_zero")
        ops._count(2)

    def update(self, logits, labels, loss=None):
        """logits [B,C] fp32 (C = 2: softmax[:,1] and argmax; C = 1: sigmoid), labels [B] int64, loss: 0-d fp32 device tensor."""
        if not (logits.is_cuda and labels.is_cuda):
            raise RuntimeError("DeviceMetrics.update needs device tensors")
        assert logits.dtype == torch.float32 and labels.dtype == torch.int64 and logits.dim() == 2
        logits = logits.contiguous()
        if loss is not None:
            assert loss.dtype == torch.float32 and loss.numel() == 1
        B, C = logits.shape
        check(_lib.load().dfu_metrics_update(_p(logits), _p(labels.contiguous()), _p(loss), B, C, _p(self.state), _p(self.loss_acc),
                                             _p(self.probs), _p(self.y_true), _p(self.y_pred), self.capacity,
                                             torch.cuda.current_stream().cuda_stream), "metrics_update")
        ops._count()

    def compute(self):
        """All metrics of extended_metrics.py:374-446 (same names) plus mean_loss and n: one kernel, ONE device->host copy."""
        check(_lib.load().dfu_metrics_finalize(_p(self.state), _p(self.loss_acc), _p(self.probs), _p(self.y_true), self.capacity,
                                               _p(self.out), torch.cuda.current_stream().cuda_stream), "metrics_finalize")
        ops._count()
        vals = self.out.cpu().tolist()
        m = dict(zip(NAMES, vals))
        for k in ("tn", "fp", "fn", "tp", "n"):
            m[k] = int(m[k])
        m["overflowed"] = bool(m["overflowed"])
        m["confusion_matrix"] = [[m["tn"], m["fp"]], [m["fn"], m["tp"]]]
        return m

    def samples(self):
        """(y_true, y_pred, y_probs) seen so far, as the reference's evaluate_* functions return them (one sync)."""
        n = min(int(self.state[0].item()), self.capacity)
        return self.y_true[:n].cpu().numpy(), self.y_pred[:n].cpu().numpy(), self.probs[:n].cpu().numpy()
