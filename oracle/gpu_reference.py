"""oracle/gpu_reference.py -- the "reference GPU build" comparator (SURVEY.md §8d last row): the reference's step as
eager PyTorch ON THE GPU, with the reference's own CUDA operators where it has them.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py's `gpu_reference` leg and tests/test_gpu_reference.py).  It is the baseline the product is
compared with, never part of the product path.

What runs: * the UNet as the torch restatement oracle/unet_oracle.py on CUDA tensors (cuDNN convolutions, cuBLAS matmuls; TF32 allowed, which
is what the reference's pinned torch 1.11 defaulted to -- SURVEY.md §0.6);  * the VAE as oracle/vae_oracle.py and the LPIPS term as
oracle/lpips_oracle.py on CUDA;  * the NGP render as the LITERAL three-pass `run()` of external/nerf/renderer_df.py:310-468 (density pass :373,
importance pass :398, colour pass :424 -- not the single-evaluation restatement) in torch on CUDA, with the grid encoder and
near_far_from_aabb executed by the REFERENCE's own CUDA kernels compiled into oracle/_ref/ (external/gridencoder/src/gridencoder.cu,
raymarching/src/raymarching.cu; oracle/build_ref.py), wrapped the way external/gridencoder/grid.py:19-88 wraps them; torch.optim.Adam.
The reference itself cannot be imported on the GPU box (/root/reference does not travel; its python needs pytorch3d, trimesh, ...), so this
is the closest executable statement of "the reference GPU build" there: same operators, same eager op sequence, same arithmetic class.
"""
from __future__ import annotations

import math
import os
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from . import build_ref
from . import ngp_oracle as no

_MODS: Dict[str, object] = {}


def _mod(name):
    if name not in _MODS:
        if not os.path.exists(build_ref.so_path(name)):
            raise RuntimeError(f'oracle/_ref/{name}.so is missing (built from /root/reference by oracle/build_ref.py in the build container)')
        _MODS[name] = build_ref.load_module(name)
    return _MODS[name]


class _RefGridEncode(torch.autograd.Function):
    """external/gridencoder/grid.py:19-88 around the reference's CUDA kernels"""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, S, H):
        m = _mod('_ref_gridencoder')
        inputs = inputs.contiguous()
        B, D = inputs.shape
        L, C = offsets.shape[0] - 1, embeddings.shape[1]
        outputs = torch.empty(L, B, C, device=inputs.device, dtype=embeddings.dtype)
        m.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, None, 1, False)
        ctx.save_for_backward(inputs, embeddings, offsets)
        ctx.dims = (B, D, C, L, S, H)
        return outputs.permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        m = _mod('_ref_gridencoder')
        inputs, embeddings, offsets = ctx.saved_tensors
        B, D, C, L, S, H = ctx.dims
        grad = grad.view(B, L, C).permute(1, 0, 2).contiguous()
        grad_embeddings = torch.zeros_like(embeddings)
        m.grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, None, None, 1, False)
        return None, grad_embeddings, None, None, None


class RefCudaField:
    """external/nerf/network_grid.py:36-88,167-208 (NeRFNetwork) as a function of a CUDA parameter dict"""

    def __init__(self, params: Dict[str, torch.Tensor], bound: float = 4.0):
        self.p, self.bound = params, bound
        geo = no.live_geometry(bound)
        dev = params['encoder.embeddings'].device
        self.offsets = torch.from_numpy(geo['offsets']).to(dev)
        self.S, self.H = float(geo['S']), int(geo['H'])

    def encode(self, x):
        return _RefGridEncode.apply(((x + self.bound) / (2 * self.bound)).reshape(-1, 3), self.p['encoder.embeddings'], self.offsets, self.S, self.H)

    def common_forward(self, x):
        h = self.encode(x)
        h = F.relu(F.linear(h, self.p['sigma_net.net.0.weight'], self.p['sigma_net.net.0.bias']))
        h = F.relu(F.linear(h, self.p['sigma_net.net.1.weight'], self.p['sigma_net.net.1.bias']))
        h = F.linear(h, self.p['sigma_net.net.2.weight'], self.p['sigma_net.net.2.bias'])
        blob = 5 * torch.exp(-(x ** 2).sum(-1) / (2 * 0.2 ** 2))
        return no._TruncExp.apply(h[..., 0] + blob), torch.sigmoid(h[..., 1:])

    def density(self, x):
        sigma, albedo = self.common_forward(x)
        return {'sigma': sigma, 'albedo': albedo}


def near_far(rays_o, rays_d, aabb, min_near):
    m = _mod('_ref_raymarching')
    N = rays_o.shape[0]
    nears, fars = torch.empty(N, device=rays_o.device), torch.empty(N, device=rays_o.device)
    m.near_far_from_aabb(rays_o.contiguous(), rays_d.contiguous(), aabb, N, float(min_near), nears, fars)
    return nears, fars


def run_three_pass(field, rays_o, rays_d, *, num_steps=64, upsample_steps=64, min_near=0.1, bg_color=0.0, perturb_noise=None, pdf_noise=None,
                   training=True, near_far_fn=near_far):
    """renderer_df.py:310-468 literally: density(xyzs) -> sample_pdf -> density(new_xyzs) -> sort / gather -> self(xyzs) again for the colours.
    Device agnostic given `field` and `near_far_fn` (CPU: pass oracle stand-ins)."""
    dev = rays_o.device
    rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    N = rays_o.shape[0]
    b = field.bound
    aabb = torch.tensor([-b, -b, -b, b, b, b], dtype=torch.float32, device=dev)
    nears, fars = near_far_fn(rays_o, rays_d, aabb, min_near)
    nears, fars = nears.unsqueeze(-1), fars.unsqueeze(-1)
    z_vals = torch.linspace(0.0, 1.0, num_steps, device=dev).unsqueeze(0).expand((N, num_steps))
    z_vals = nears + (fars - nears) * z_vals
    sample_dist = (fars - nears) / num_steps
    if perturb_noise is not None:
        z_vals = z_vals + (perturb_noise - 0.5) * sample_dist
    xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z_vals.unsqueeze(-1)
    xyzs = torch.min(torch.max(xyzs, aabb[:3]), aabb[3:])
    density_outputs = {k: v.view(N, num_steps, -1) for k, v in field.density(xyzs.reshape(-1, 3)).items()}            # pass 1 (:373)
    with torch.no_grad():
        deltas = z_vals[..., 1:] - z_vals[..., :-1]
        deltas = torch.cat([deltas, sample_dist * torch.ones_like(deltas[..., :1])], dim=-1)
        alphas = 1 - torch.exp(-deltas * density_outputs['sigma'].squeeze(-1))
        alphas_shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
        weights = alphas * torch.cumprod(alphas_shifted, dim=-1)[..., :-1]
        z_vals_mid = (z_vals[..., :-1] + 0.5 * deltas[..., :-1])
        new_z_vals = no.sample_pdf(z_vals_mid, weights[:, 1:-1], upsample_steps, det=not training,
                                   u=pdf_noise if training else torch.linspace(0.5 / upsample_steps, 1 - 0.5 / upsample_steps, upsample_steps, device=dev).expand(N, upsample_steps)).detach()
        new_xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * new_z_vals.unsqueeze(-1)
        new_xyzs = torch.min(torch.max(new_xyzs, aabb[:3]), aabb[3:])
    new_density_outputs = {k: v.view(N, upsample_steps, -1) for k, v in field.density(new_xyzs.reshape(-1, 3)).items()}  # pass 2 (:398)
    z_vals = torch.cat([z_vals, new_z_vals], dim=1)
    z_vals, z_index = torch.sort(z_vals, dim=1)
    xyzs = torch.cat([xyzs, new_xyzs], dim=1)
    xyzs = torch.gather(xyzs, dim=1, index=z_index.unsqueeze(-1).expand_as(xyzs))
    for k in density_outputs:
        tmp = torch.cat([density_outputs[k], new_density_outputs[k]], dim=1)
        density_outputs[k] = torch.gather(tmp, dim=1, index=z_index.unsqueeze(-1).expand_as(tmp))
    deltas = z_vals[..., 1:] - z_vals[..., :-1]
    deltas = torch.cat([deltas, sample_dist * torch.ones_like(deltas[..., :1])], dim=-1)
    alphas = 1 - torch.exp(-deltas * density_outputs['sigma'].squeeze(-1))
    alphas_shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
    weights = alphas * torch.cumprod(alphas_shifted, dim=-1)[..., :-1]
    _, rgbs = field.common_forward(xyzs.reshape(-1, 3))                                                                   # pass 3 (:424), albedo shading
    rgbs = rgbs.view(N, -1, 3)
    weights_sum = weights.sum(dim=-1)
    ori_z_vals = ((z_vals - nears) / (fars - nears)).clamp(0, 1)
    depth = torch.sum(weights * ori_z_vals, dim=-1)
    image = torch.sum(weights.unsqueeze(-1) * rgbs, dim=-2)
    image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
    return dict(image=image, depth=depth, weights_sum=weights_sum)


def _time(fn, warm=2, iters=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def time_step_components(device, mean_unet_calls: float) -> dict:
    """milliseconds of each component of one SDS iteration on `device`, assembled into a step:
    2 x (render forward + loss backward) + (n+1) UNet evaluations + VAE encode/decode + LPIPS value/gradient + 2 Adam updates."""
    from . import lpips_oracle as lo, unet_oracle as uo, vae_oracle as vo
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = True          # torch 1.11's defaults
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        with torch.cuda.device(device):
            cfg = uo.FULL
            sd = {k: v.to(device) for k, v in uo.make_params(cfg, seed=0).items()}
            x, c = torch.randn(1, 4, 32, 32, device=device), torch.randn(1, 256, 32, 32, device=device)
            ls = uo.alpha_cosine_log_snr(torch.tensor([0.3], device=device))
            with torch.no_grad():
                t_unet = _time(lambda: uo.unet_forward(sd, cfg, x, ls, c), warm=3, iters=10)
            vae = vo.TorchVAE(vo.make_params(seed=0)).to(device)
            img = torch.rand(1, 3, 256, 256, device=device)
            z = torch.randn(1, 4, 32, 32, device=device)
            t_vae = _time(lambda: (vae.encode(img * 2 - 1).mode(), vae.decode(z)))
            lp = {k: v.to(device) for k, v in lo.make_params(0).items()}

            def lpips_fb():
                a = img.clone().requires_grad_(True)
                lo.lpips(lp, 2 * a - 1, 2 * img.flip(-1) - 1).sum().backward()
            t_lpips = _time(lpips_fb)
            params = {k: v.to(device).requires_grad_(True) for k, v in no.make_field_params(seed=0).items()}
            field = RefCudaField(params)
            opt = torch.optim.Adam([{'params': [params['encoder.embeddings']], 'lr': 5e-3}, {'params': [params[k] for k in no.PARAM_KEYS[1:]], 'lr': 5e-4}])
            ro, rd = (torch.from_numpy(a).to(device) for a in no.camera_rays(no.circle_cameras(64)[3], 128, 128))
            N = ro.shape[0]

            def render_fb():
                r = run_three_pass(field, ro, rd, perturb_noise=torch.rand(N, 64, device=device), pdf_noise=torch.rand(N, 64, device=device))
                opt.zero_grad()
                (r['image'].mean() + r['weights_sum'].mean()).backward()
                opt.step()
            t_render = _time(render_fb)
        step = 2 * t_render + mean_unet_calls * t_unet + t_vae + t_lpips
        return {'kind': 'eager PyTorch restatement on the same GPU (cuDNN/cuBLAS, TF32 on) + the reference\'s own CUDA operators (oracle/_ref) for the grid encoder / '
                        'near_far; literal three-pass run(); components timed separately and summed (no CUDA graphs, as in the reference)',
                'ms_per_step': round(step, 3), 'steps_per_s': round(1e3 / step, 4), 'unet_eval_ms': round(t_unet, 3), 'unet_evals_per_step': round(mean_unet_calls, 2),
                'render_fwd_bwd_adam_ms': round(t_render, 3), 'vae_encode_decode_ms': round(t_vae, 3), 'lpips_fwd_bwd_ms': round(t_lpips, 3),
                'tf32': True}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
