"""Grad-CAM (a11) with the observable behaviour of the reference's class (notebooks/grad_cam_visualization.py:327-429),
quirks included:

  * the target is the LAST module of `model.named_modules()` whose name contains one of the target strings
    (`layer4` -> `layer4.2.relu`; `blocks` -> `blocks.11.mlp.fc2`);
  * the activation is what that module returned on its LAST forward call of the pass, the gradient is the one that
    reached any of its outputs LAST during backward (for the shared `layer4.2.relu` that is the 512-channel gradient of
    its first call against the 2048-channel activation of its third);
  * the score is `output[0, 0]` whatever `class_idx` says; `min(C_A, C_G)` channels are reduced;
  * activations that are not 4-D (the ViT branch) fall back to the input-gradient saliency map.

The plumbing is not the reference's: only the ONE resolved target module is hooked (the reference hooks all 28 matching
modules and keeps two name-keyed dictionaries), so only the Bottleneck that owns it leaves the fused path.  The backward
pass computes only what the map needs: parameters are frozen for the pass (no weight gradients -- the reference computes
and discards them), and `torch.autograd.grad` asks for the gradient AT the target's outputs (4-D targets) or at the input
(saliency fallback), so a ResNet map backpropagates through `layer4.2` only instead of the whole branch.  Which gradient is
"the" gradient is decided as the reference's hooks decide it: the one that arrives last, i.e. the target's FIRST call.
The reduction runs in one kernel per sample (dfu_gradcam / dfu_saliency) instead of a Python loop over channels
(:421-422), and `generate_cam_batch` produces the maps of a whole batch from ONE forward + backward pass (eval-mode
samples are independent, so the gradient of sum_i output[i, 0] is every sample's own gradient)."""
import torch

from . import ops


class GradCAM:
    def __init__(self, model, target_layers):
        self.model = model
        self.target_layers = list(target_layers) if isinstance(target_layers, (list, tuple)) else [target_layers]
        self.target_name, self.target_module = None, None
        for name, module in model.named_modules():
            if any(t in name for t in self.target_layers):
                self.target_name, self.target_module = name, module          # the last match wins
        self._act = None        # latest forward output of the target
        self._outs = []         # outputs of every call of the target in the current pass
        self._handle = None
        if self.target_module is not None:
            self._handle = self.target_module.register_forward_hook(self._on_forward)

    # -- hooks ---------------------------------------------------------------------------------------------
    def _on_forward(self, module, inputs, output):
        self._act = output                              # the LAST call's output is the activation
        if isinstance(output, torch.Tensor) and output.requires_grad:
            self._outs.append(output)                   # every call's output may receive a gradient

    def close(self):
        """The reference never removes its hooks (a new GradCAM per sample piles them up); this does."""
        if self._handle is not None:
            self._handle.remove()
            self._handle = None

    # -- one pass --------------------------------------------------------------------------------------------
    def _pass(self, inputs):
        """Forward of the branch and the backward the map needs.  Returns (x, activation, gradient at the target | None,
        input gradient | None)."""
        self.model.eval()
        self._act, self._outs = None, []
        x = inputs.clone().detach().requires_grad_(True)
        params = [p for p in self.model.parameters() if p.requires_grad]
        for p in params:
            p.requires_grad_(False)                     # no weight gradients for a heat map
        try:
            with torch.enable_grad():
                out = self.model(x)
                score = out[:, 0].sum()                 # batch 1: exactly output[0, 0] (grad_cam_visualization.py:385-386)
                act = self._act
                if act is not None and act.ndim == 4 and self._outs:
                    grads = torch.autograd.grad(score, self._outs, allow_unused=True)
                    # backward visits the target's calls in reverse order, the reference keeps the gradient written last:
                    # that of the earliest call that receives one
                    grad = next((g for g in grads if g is not None), None)
                    return x, act, grad, None
                (gx,) = torch.autograd.grad(score, [x], allow_unused=True)
                return x, act, None, gx
        finally:
            for p in params:
                p.requires_grad_(True)

    def _maps(self, x, act, grad, gx, n):
        """List of n numpy maps."""
        if act is None:
            return [None] * n
        hw_in = tuple(x.shape[2:])
        if act.ndim != 4 or grad is None or grad.ndim != 4:
            if gx is None:
                return [torch.zeros(hw_in).numpy() for _ in range(n)]
            g = gx.detach()
            return [ops.saliency(g[i]).cpu().numpy() for i in range(n)]
        a = ops.as_nhwc(act.detach())
        g = ops.as_nhwc(grad.detach())
        _, h, w, ca = a.shape
        cg = g.shape[3]
        if g.shape[1] != h or g.shape[2] != w:
            raise RuntimeError("Grad-CAM activation/gradient spatial sizes differ: %s vs %s" % (a.shape, g.shape))
        nc = min(ca, cg)
        cams = [ops.gradcam_reduce(a[i].reshape(h * w, ca), g[i].reshape(h * w, cg), nc).view(h, w) for i in range(n)]
        return [c.numpy() for c in torch.stack(cams).cpu()]

    def generate_cam(self, input_tensor, class_idx=None):
        """input (1,C,H,W) -> numpy heat map: (h,w) of the target layer for 4-D activations, else (H,W) saliency.
        `class_idx` is accepted and ignored, as in the reference."""
        if self.target_module is None:
            return None
        return self._maps(*self._pass(input_tensor), 1)[0]

    def generate_cam_batch(self, inputs):
        """inputs (B,C,H,W) -> list of B maps, each equal to generate_cam(inputs[i:i+1]), from one forward + backward."""
        if self.target_module is None:
            return [None] * inputs.shape[0]
        return self._maps(*self._pass(inputs), inputs.shape[0])
