"""GridEncoder / grid_encode -- host-side mirror of external/gridencoder/grid.py (reference).

Same class name, constructor keywords, parameter/buffer names (``embeddings``, ``offsets``) and
forward signature ``forward(inputs, bound=1)``, so NeRF checkpoints round-trip and
external/ngp_encoder.py:69-71 can construct it unchanged.  The arithmetic runs in
libsparsefusion_b200.so through the `_gridencoder` operator boundary.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _gridencoder as _backend

_gridtype_to_id = {'hash': 0, 'tiled': 1}


class _grid_encode(Function):
    """autograd wrapper: grid.py:19-88 (reference).  inputs [B,D] in [0,1] -> [B, L*C]."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False):
        inputs = inputs.contiguous().float()
        embeddings = embeddings.float()
        B, D = inputs.shape
        L, C = offsets.shape[0] - 1, embeddings.shape[1]
        S, H = float(np.log2(per_level_scale)), int(base_resolution)
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=torch.float32)
        dy_dx = torch.empty(B, L * D * C, device=inputs.device, dtype=torch.float32) if calc_grad_inputs else None
        _backend.grid_encode_forward(inputs, embeddings.contiguous(), offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.dims = (B, D, C, L, S, H, gridtype, align_corners)
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        B, D, C, L, S, H, gridtype, align_corners = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        grad_embeddings = torch.zeros_like(embeddings)
        grad_inputs = torch.zeros_like(inputs) if dy_dx is not None else None
        _backend.grid_encode_backward(grad, inputs, embeddings.contiguous(), offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs,
                                      gridtype, align_corners)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None


grid_encode = _grid_encode.apply


def level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners):
    """row offsets of each level: grid.py:110-121 (reference)"""
    cap = 2 ** log2_hashmap_size
    offs, total = [], 0
    for lvl in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** lvl))
        rows = min(cap, (res if align_corners else res + 1) ** input_dim)
        rows = int(np.ceil(rows / 8) * 8)
        offs.append(total)
        total += rows
    offs.append(total)
    return np.asarray(offs, dtype=np.int32)


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype='hash', align_corners=False):
        super().__init__()
        if desired_resolution is not None:
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.log2_hashmap_size, self.base_resolution = per_level_scale, log2_hashmap_size, base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype, self.gridtype_id, self.align_corners = gridtype, _gridtype_to_id[gridtype], align_corners
        self.max_params = 2 ** log2_hashmap_size
        offsets = torch.from_numpy(level_offsets(input_dim, num_levels, per_level_scale, base_resolution, log2_hashmap_size, align_corners))
        self.register_buffer('offsets', offsets)
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)

    def __repr__(self):
        return (f'GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} '
                f'resolution={self.base_resolution} -> {int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} '
                f'per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} gridtype={self.gridtype} '
                f'align_corners={self.align_corners}')

    def forward(self, inputs, bound=1):
        inputs = (inputs + bound) / (2 * bound)
        prefix = list(inputs.shape[:-1])
        inputs = inputs.view(-1, self.input_dim)
        out = grid_encode(inputs, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution, inputs.requires_grad,
                          self.gridtype_id, self.align_corners)
        return out.view(prefix + [self.output_dim])
