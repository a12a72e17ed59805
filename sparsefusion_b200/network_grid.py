"""NeRFNetwork -- host-side mirror of external/nerf/network_grid.py (reference): the Instant-NGP scene that
SparseFusion optimises (sparsefusion/distillation.py:164-165).

Same constructor (``NeRFNetwork(opt)``), parameter names (``encoder.embeddings``, ``encoder.offsets``,
``sigma_net.net.{0,1,2}.{weight,bias}``) so checkpoints round-trip (distillation.py:495-496), same methods
(``common_forward``, ``density``, ``forward(x, d, l, ratio, shading)``, ``get_params(lr)``).  The field query
(grid encode + MLP + activations) is ONE fused kernel with a fused backward (csrc/ngp_field.cu) instead of the
reference's kernel_grid + permute + three cuBLAS GEMMs + elementwise launches.  Only shading='albedo' exists
(the one SparseFusion uses); normal-based shadings of torch-ngp are out of scope.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import _lib as lib
from .gridencoder import GridEncoder
from .renderer_df import NeRFRenderer


def get_default_torch_ngp_opt():
    """sparsefusion/distillation.py:500-526 (the options the render kernels are parameterised by)"""
    return SimpleNamespace(cuda_ray=False, max_steps=256, num_steps=64, upsample_steps=64, update_extra_interval=16, max_ray_batch=4096,
                           albedo_iters=1000, bg_radius=0, density_thresh=10, fp16=True, backbone='grid', w=128, h=128, hw_scale=2, bound=4,
                           min_near=0.1, dt_gamma=0, lambda_entropy=1e-4, lambda_opacity=0, lambda_orient=1e-2, lambda_smooth=0)


class MLP(nn.Module):
    """parameter container with the reference's layout (network_grid.py:14-33); evaluated inside the fused field kernel"""

    def __init__(self, dim_in, dim_out, dim_hidden, num_layers, bias=True):
        super().__init__()
        self.dim_in, self.dim_out, self.dim_hidden, self.num_layers = dim_in, dim_out, dim_hidden, num_layers
        self.net = nn.ModuleList([nn.Linear(dim_in if l == 0 else dim_hidden, dim_out if l == num_layers - 1 else dim_hidden, bias=bias)
                                  for l in range(num_layers)])


class _FieldFn(torch.autograd.Function):
    """sigma, albedo = common_forward(x)   (network_grid.py:77-88) for explicit points x [B,3]"""

    @staticmethod
    def forward(ctx, net, x, emb, w0, b0, w1, b1, w2, b2):
        enc = net.encoder
        B = x.shape[0]
        sigma = torch.empty(B, device=x.device)
        rgb = torch.empty(B, 3, device=x.device)
        f = lib.fptr
        ctx.geo = (float(math.log2(enc.per_level_scale)), int(enc.base_resolution), float(net.bound))
        lib.call('sfb_ngp_field_forward', f(x), None, None, None, 0, B, f(emb), lib.iptr(enc.offsets), *ctx.geo, f(w0), f(b0), f(w1), f(b1), f(w2),
                 f(b2), f(sigma), f(rgb), lib.stream())
        ctx.save_for_backward(x, emb, w0, b0, w1, b1, w2, b2)
        ctx.offsets = enc.offsets
        return sigma, rgb

    @staticmethod
    def backward(ctx, g_sigma, g_rgb):
        x, emb, w0, b0, w1, b1, w2, b2 = ctx.saved_tensors
        B = x.shape[0]
        f = lib.fptr
        g_sigma = torch.zeros(B, device=x.device) if g_sigma is None else g_sigma.contiguous()
        g_rgb = None if g_rgb is None else g_rgb.contiguous()
        tape = torch.empty(lib.load().sfb_ngp_field_tape_floats(B), device=x.device)
        g_emb = torch.zeros_like(emb)
        gw0, gb0, gw1, gb1, gw2, gb2 = (torch.zeros_like(t) for t in (w0, b0, w1, b1, w2, b2))
        lib.call('sfb_ngp_field_backward', f(x), None, None, None, 0, B, f(emb), lib.iptr(ctx.offsets), *ctx.geo, f(w0), f(b0), f(w1), f(b1), f(w2),
                 f(b2), f(g_sigma), f(g_rgb), f(g_emb), f(gw0), f(gb0), f(gw1), f(gb1), f(gw2), f(gb2), f(tape), lib.stream())
        return None, None, g_emb, gw0, gb0, gw1, gb1, gw2, gb2


class NeRFNetwork(NeRFRenderer):
    def __init__(self, opt, num_layers=3, hidden_dim=64, num_layers_bg=2, hidden_dim_bg=64):
        super().__init__(opt)
        if num_layers != 3 or hidden_dim != 64:
            raise NotImplementedError('the fused field kernel is built for the SparseFusion MLP (3 layers, 64 hidden)')
        if self.bg_radius > 0:
            raise NotImplementedError('bg_radius > 0 (background network) is not used by SparseFusion (distillation.py:512)')
        self.num_layers, self.hidden_dim = num_layers, hidden_dim
        # get_encoder('tiledgrid', input_dim=3, log2_hashmap_size=16, desired_resolution=2048*bound)  (network_grid.py:50, ngp_encoder.py:69-71)
        self.encoder = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16,
                                   desired_resolution=2048 * self.bound, gridtype='tiled', align_corners=False)
        self.in_dim = self.encoder.output_dim
        self.sigma_net = MLP(self.in_dim, 4, hidden_dim, num_layers, bias=True)
        self.bg_net = None

    def _field_params(self):
        n = self.sigma_net.net
        return (self.encoder.embeddings, n[0].weight, n[0].bias, n[1].weight, n[1].bias, n[2].weight, n[2].bias)

    def gaussian(self, x):
        d = (x ** 2).sum(-1)
        return 5 * torch.exp(-d / (2 * 0.2 ** 2))

    def common_forward(self, x):
        """x: [N, 3] in [-bound, bound] -> sigma [N], albedo [N, 3]"""
        if not x.is_cuda:
            raise RuntimeError('NeRFNetwork.common_forward needs CUDA tensors (no CPU fallback)')
        x = x.contiguous().view(-1, 3).float()
        with torch.cuda.device(x.device):
            return _FieldFn.apply(self, x, *self._field_params())

    def forward(self, x, d, l=None, ratio=1, shading='textureless'):
        if shading != 'albedo':
            raise NotImplementedError("only shading='albedo' is part of SparseFusion's path (distillation.py:209)")
        sigma, color = self.common_forward(x)
        return sigma, color, None

    def density(self, x):
        sigma, albedo = self.common_forward(x)
        return {'sigma': sigma, 'albedo': albedo}

    def get_params(self, lr):
        return [{'params': self.encoder.parameters(), 'lr': lr * 10}, {'params': self.sigma_net.parameters(), 'lr': lr}]
