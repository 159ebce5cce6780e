"""Grad-CAM (a11) with the same observable behaviour as the reference class
(notebooks/grad_cam_visualization.py:327-429), including its quirks: forward hooks on every
module whose name contains a target string, score = output[0, 0] (class_idx is ignored), the LAST matching module
name supplies activations and gradients, channel count min(C_A, C_G), and an input-gradient saliency fallback for
non-4D activations (the ViT branch).  The reduction itself runs in one kernel (dfu_gradcam / dfu_saliency) instead
of the reference's per-channel Python loop (:421-422)."""
import torch

from . import ops


class GradCAM:
    def __init__(self, model, target_layers):
        self.model = model
        self.target_layers = target_layers if isinstance(target_layers, list) else [target_layers]
        self.activations = {}
        self.gradients = {}
        self._handles = []
        self._register_hooks()

    def _register_hooks(self):
        def get_activation(name):
            def hook(module, inputs, output):
                self.activations[name] = output
                if isinstance(output, torch.Tensor) and output.requires_grad:
                    def save_grad(grad):
                        self.gradients[name] = grad
                    output.register_hook(save_grad)
            return hook

        for name, module in self.model.named_modules():
            if any(layer in name for layer in self.target_layers):
                self._handles.append(module.register_forward_hook(get_activation(name)))

    def close(self):
        """The reference never removes its hooks (a new GradCAM per sample piles them up); we offer a way out."""
        for h in self._handles:
            h.remove()
        self._handles = []

    def generate_cam(self, input_tensor, class_idx=None):
        """input (1,C,H,W) -> numpy heat map: (h,w) of the target layer for 4-D activations, else (H,W) saliency."""
        self.model.eval()
        x = input_tensor.clone().detach().requires_grad_(True)
        with torch.enable_grad():
            output = self.model(x)
            self.model.zero_grad()
            score = output[0, 0]
            score.backward()

        layer_name = None
        for name, _ in self.model.named_modules():
            if any(layer in name for layer in self.target_layers):
                layer_name = name
        if layer_name is None:
            return None
        act = self.activations[layer_name]
        grad = self.gradients.get(layer_name, None)

        if act.ndim != 4 or grad is None or grad.ndim != 4:
            if x.grad is None:
                return torch.zeros(input_tensor.shape[2:]).numpy()
            return ops.saliency(x.grad.detach()[0]).cpu().numpy()

        a = ops.as_nhwc(act.detach())[0]
        g = ops.as_nhwc(grad.detach())[0]
        h, w, ca = a.shape
        cg = g.shape[2]
        if g.shape[0] != h or g.shape[1] != w:
            raise RuntimeError("Grad-CAM activation/gradient spatial sizes differ: %s vs %s" % (a.shape, g.shape))
        cam = ops.gradcam_reduce(a.reshape(h * w, ca), g.reshape(h * w, cg), min(ca, cg))
        return cam.view(h, w).cpu().numpy()
