"""Synthetic scenes for benches and smoke runs (no dataset, no pytorch3d): BASELINE.json config 3's "synthetic
hydrant-shaped cameras" -- an object-centric fly-around as utils/camera_utils.py:190-259 (get_interpolated_path) produces,
rays on the 128x128 pixel-centre grid of utils/render_utils.py:40-65 (NDC from 1-1/W to -1+1/W, directions not
normalised, as pytorch3d's GridRaysampler emits them).
"""
from __future__ import annotations

import math

import numpy as np
import torch


def circle_cameras(n_views: int, radius: float = 5.0, elevation_deg: float = 15.0):
    """camera centres on a circle at the given elevation, looking at the origin, +y up -> list of (centre [3], R [3,3] rows right/up/fwd)"""
    el = math.radians(elevation_deg)
    cams = []
    for i in range(n_views):
        az = 2 * math.pi * i / n_views
        c = np.array([radius * math.cos(el) * math.sin(az), radius * math.sin(el), radius * math.cos(el) * math.cos(az)])
        fwd = -c / np.linalg.norm(c)
        right = np.cross(np.array([0.0, 1.0, 0.0]), fwd)
        right /= np.linalg.norm(right)
        up = np.cross(fwd, right)
        cams.append((c.astype(np.float32), np.stack([right, up, fwd]).astype(np.float32)))
    return cams


def camera_rays(cam, H: int, W: int, focal_ndc: float = 4.0):
    c, R = cam
    xs = np.linspace(1 - 1 / W, -1 + 1 / W, W, dtype=np.float32)
    ys = np.linspace(1 - 1 / H, -1 + 1 / H, H, dtype=np.float32)
    yy, xx = np.meshgrid(ys, xs, indexing='ij')
    d = np.stack([xx / focal_ndc, yy / focal_ndc, np.ones_like(xx)], axis=-1).reshape(-1, 3) @ R
    return np.ascontiguousarray(np.broadcast_to(c, d.shape), np.float32), np.ascontiguousarray(d, np.float32)


def synthetic_scene(n_input=2, n_target=64, image_size=256, latent=32, feat_ch=256, render_hw=128, seed=0, radius=5.0):
    """dict of CPU tensors with the field names of distillation.SceneCache"""
    rng = np.random.default_rng(seed)
    cams = circle_cameras(n_target, radius=radius)
    yy, xx = np.mgrid[0:image_size, 0:image_size]
    disc = (((yy - image_size / 2) ** 2 + (xx - image_size / 2) ** 2) < (0.35 * image_size) ** 2).astype(np.float32)

    def rays(idx):
        o, d = zip(*[camera_rays(cams[i], render_hw, render_hw) for i in idx])
        return torch.from_numpy(np.stack(o)), torch.from_numpy(np.stack(d))
    in_idx = [int(i * n_target / n_input) for i in range(n_input)]
    iro, ird = rays(in_idx)
    tro, trd = rays(range(n_target))
    return dict(
        input_rgb=torch.from_numpy(rng.random((n_input, 3, image_size, image_size), dtype=np.float32) * disc),
        input_mask=torch.from_numpy(np.broadcast_to(disc, (n_input, 1, image_size, image_size)).copy()),
        input_rays_o=iro, input_rays_d=ird,
        target_features=torch.from_numpy(rng.standard_normal((n_target, feat_ch, latent, latent), dtype=np.float32)),
        target_eft_image=torch.from_numpy(rng.random((n_target, 3, image_size, image_size), dtype=np.float32) * disc),
        target_rays_o=tro, target_rays_d=trd)
