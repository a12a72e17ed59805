"""CustomImplicitRenderer -- host-side mirror of utils/eft_renderer.py:17-166 (reference): the pytorch3d-style implicit renderer the distillation
loop drives the EFT through when it builds the per-view feature cache (sparsefusion/distillation.py:85-86, :103-110 via
utils/render_utils.py:init_light_field_renderer).

Pure orchestration, same constructor / ``forward(cameras, volumetric_function, **kwargs) -> (images, ray_bundle, reg)`` contract:
``raysampler(cameras=..., volumetric_function=..., **kwargs)`` -> ray bundle; ``volumetric_function(ray_bundle=..., cameras=..., **kwargs)``
-> (densities, features, reg); ``raymarcher(rays_densities=..., rays_features=..., ray_bundle=..., **kwargs)`` -> images.  No pytorch3d import is
needed for that (the reference only imports pytorch3d types for annotations), so this class also works with the ray samplers / EFT of
sparsefusion_b200 itself.
"""
from __future__ import annotations

from typing import Callable

import torch


class CustomImplicitRenderer(torch.nn.Module):
    def __init__(self, raysampler: Callable, raymarcher: Callable, reg=None) -> None:
        super().__init__()
        if not callable(raysampler):
            raise ValueError('"raysampler" has to be a "Callable" object.')
        if not callable(raymarcher):
            raise ValueError('"raymarcher" has to be a "Callable" object.')
        self.raysampler = raysampler
        self.raymarcher = raymarcher
        self.reg = reg

    def forward(self, cameras, volumetric_function: Callable, **kwargs):
        if not callable(volumetric_function):
            raise ValueError('"volumetric_function" has to be a "Callable" object.')
        ray_bundle = self.raysampler(cameras=cameras, volumetric_function=volumetric_function, **kwargs)           # eft_renderer.py:133-135
        rays_densities, rays_features, reg_term = volumetric_function(ray_bundle=ray_bundle, cameras=cameras, **kwargs)   # :144-151
        images = self.raymarcher(rays_densities=rays_densities, rays_features=rays_features, ray_bundle=ray_bundle, **kwargs)   # :156-161
        return (images, ray_bundle, reg_term) if self.reg is not None else (images, ray_bundle, 0)                 # :164-167
