"""oracle/ngp_oracle.py -- CPU restatement of the reference's torch-ngp render path.

TEST INFRASTRUCTURE ONLY (see oracle/ngp_oracle.c): only tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs import this module.

Layers restated here (citations relative to the reference repo):

* operator level -- ctypes bindings to oracle/ngp_oracle.c, one per pybind function of
  ``_gridencoder`` (external/gridencoder/src/bindings.cpp:5-8) and ``_raymarching``
  (raymarching/src/bindings.cpp:5-18).
* ``GridEncoder`` geometry and autograd wrapper -- external/gridencoder/grid.py:19-154.
* ``NeRFNetwork.common_forward / density / forward(shading='albedo')`` --
  external/nerf/network_grid.py:14-33, 69-88, 167-208; ``trunc_exp`` external/ngp_activation.py:10-21.
* ``sample_pdf`` and ``NeRFRenderer.run`` -- external/nerf/renderer_df.py:15-49, 310-468.
* ``NeRFRenderer.run_cuda`` (train + eval) and ``update_extra_state`` -- renderer_df.py:471-640
  with the wrappers of raymarching/raymarching.py:161-373.

Every random draw of the reference (``torch.rand`` for perturb / sample_pdf / march noise /
density-grid jitter) is an explicit argument, so the CUDA path and the restatement can be run on
identical noise.
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, '_build')
_SO = os.path.join(_BUILD, 'libngp_oracle.so')
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 -ffp-contract=off -fopenmp oracle/ngp_oracle.c -> oracle/_build/libngp_oracle.so"""
    src = os.path.join(_HERE, 'ngp_oracle.c')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(_BUILD, exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-std=c99', '-ffp-contract=off', '-fopenmp', '-shared', '-fPIC',
                               '-o', _SO, src, '-lm'])
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_grid_level_scale.restype = ctypes.c_float
        _lib.oracle_grid_level_scale.argtypes = [ctypes.c_uint32, ctypes.c_float, ctypes.c_uint32]
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


u32, f32c = ctypes.c_uint32, ctypes.c_float

# ---------------------------------------------------------------------------------------------
# grid geometry (grid.py:96-124)
# ---------------------------------------------------------------------------------------------
GRIDTYPE = {'hash': 0, 'tiled': 1}


def grid_geometry(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2.0, base_resolution=16,
                  log2_hashmap_size=19, desired_resolution=None, align_corners=False):
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        n = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), float(per_level_scale)


def live_geometry(bound: float = 4.0):
    """network_grid.py:50 -> get_encoder('tiledgrid', input_dim=3, log2_hashmap_size=16,
    desired_resolution=2048*bound) with ngp_encoder.py:69-71 defaults (L=16, C=2, H=16)."""
    offsets, pls = grid_geometry(3, 16, 2, 2.0, 16, 16, 2048 * bound, False)
    return dict(offsets=offsets, per_level_scale=pls, S=float(np.log2(pls)), H=16, D=3, C=2, L=16,
                gridtype=GRIDTYPE['tiled'], align_corners=False)


def level_scales_host(L, S, H):
    return np.array([lib().oracle_grid_level_scale(l, S, H) for l in range(L)], dtype=np.float32)


# ---------------------------------------------------------------------------------------------
# operator level (numpy in / numpy out)
# ---------------------------------------------------------------------------------------------
def grid_encode_forward(inputs, embeddings, offsets, S, H, gridtype=1, align_corners=False, calc_dy_dx=False,
                        level_scales=None, want_rows=False):
    inputs, embeddings, offsets = _f32(inputs), _f32(embeddings), _i32(offsets)
    B, D = inputs.shape
    C, L = embeddings.shape[1], offsets.shape[0] - 1
    out = np.empty((L, B, C), np.float32)
    dy = np.empty((B, L * D * C), np.float32) if calc_dy_dx else None
    rows = np.empty((L, B, 1 << D), np.int32) if want_rows else None
    ls = None if level_scales is None else _f32(level_scales)
    lib().oracle_grid_encode_forward(_p(inputs), _p(embeddings), _p(offsets), _p(out), u32(B), u32(D), u32(C), u32(L),
                                     f32c(S), u32(H), _p(dy), u32(gridtype), ctypes.c_int(int(align_corners)), _p(ls), _p(rows))
    return out, dy, rows


def grid_encode_backward(grad, inputs, offsets, n_rows, S, H, gridtype=1, align_corners=False, dy_dx=None, level_scales=None):
    grad, inputs, offsets = _f32(grad), _f32(inputs), _i32(offsets)
    L, B, C = grad.shape
    D = inputs.shape[1]
    ge = np.zeros((n_rows, C), np.float32)
    gi = np.zeros((B, D), np.float32) if dy_dx is not None else None
    dy = None if dy_dx is None else _f32(dy_dx)
    ls = None if level_scales is None else _f32(level_scales)
    lib().oracle_grid_encode_backward(_p(grad), _p(inputs), _p(offsets), _p(ge), u32(B), u32(D), u32(C), u32(L), f32c(S),
                                      u32(H), _p(dy), _p(gi), u32(gridtype), ctypes.c_int(int(align_corners)), _p(ls))
    return ge, gi


def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().oracle_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), u32(N), f32c(min_near), _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().oracle_sph_from_ray(_p(rays_o), _p(rays_d), f32c(radius), u32(N), _p(coords))
    return coords


def morton3D(coords):
    coords = _i32(coords)
    out = np.empty(coords.shape[0], np.int32)
    lib().oracle_morton3D(_p(coords), u32(coords.shape[0]), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i32(indices)
    out = np.empty((indices.shape[0], 3), np.int32)
    lib().oracle_morton3D_invert(_p(indices), u32(indices.shape[0]), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f32(grid)
    N = grid.size // 8
    out = np.empty(N, np.uint8)
    lib().oracle_packbits(_p(grid), u32(N), f32c(thresh), _p(out))
    return out


def march_rays_train(rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises, dt_gamma=0.0, max_steps=1024, M=None):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    M = N * max_steps if M is None else M
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    rays, counter = np.empty((N, 3), np.int32), np.zeros(2, np.int32)
    bitfield = np.ascontiguousarray(bitfield, np.uint8)
    lib().oracle_march_rays_train(_p(rays_o), _p(rays_d), _p(bitfield), f32c(bound), f32c(dt_gamma), u32(max_steps), u32(N),
                                  u32(C), u32(H), u32(M), _p(_f32(nears)), _p(_f32(fars)), _p(xyzs), _p(dirs), _p(deltas),
                                  _p(rays), _p(counter), _p(_f32(noises)))
    return xyzs, dirs, deltas, rays, counter


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    ws, depth, image = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, 3), np.float32)
    lib().oracle_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), u32(M), u32(N), f32c(T_thresh),
                                              _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, ws, image, T_thresh=1e-4):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    M, N = sigmas.shape[0], rays.shape[0]
    gs, gc = np.zeros(M, np.float32), np.zeros((M, 3), np.float32)
    lib().oracle_composite_rays_train_backward(_p(_f32(grad_ws)), _p(_f32(grad_image)), _p(sigmas), _p(rgbs), _p(deltas),
                                               _p(rays), _p(_f32(ws)), _p(_f32(image)), u32(M), u32(N), f32c(T_thresh), _p(gs), _p(gc))
    return gs, gc


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises,
               dt_gamma=0.0, max_steps=1024, align=-1):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    M = n_alive * n_step
    if align > 0:
        M += align - (M % align)
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    bitfield = np.ascontiguousarray(bitfield, np.uint8)
    lib().oracle_march_rays(u32(n_alive), u32(n_step), _p(_i32(rays_alive)), _p(_f32(rays_t)), _p(rays_o), _p(rays_d),
                            f32c(bound), f32c(dt_gamma), u32(max_steps), u32(C), u32(H), _p(bitfield), _p(_f32(nears)),
                            _p(_f32(fars)), _p(xyzs), _p(dirs), _p(deltas), _p(_f32(noises)))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """in-place on rays_alive / rays_t / weights_sum / depth / image (numpy arrays of the right dtype)"""
    for a, dt in ((rays_alive, np.int32), (rays_t, np.float32), (weights_sum, np.float32), (depth, np.float32), (image, np.float32)):
        assert a.dtype == dt and a.flags['C_CONTIGUOUS']
    lib().oracle_composite_rays(u32(n_alive), u32(n_step), f32c(T_thresh), _p(rays_alive), _p(rays_t), _p(_f32(sigmas)),
                                _p(_f32(rgbs)), _p(_f32(deltas)), _p(weights_sum), _p(depth), _p(image))


# ---------------------------------------------------------------------------------------------
# autograd level: GridEncoder + NeRFNetwork field
# ---------------------------------------------------------------------------------------------
class _GridEncode(torch.autograd.Function):
    """grid.py:19-88 on CPU tensors (inputs are not differentiated on this path: grid.py:149)"""

    @staticmethod
    def forward(ctx, inputs, embeddings, geo, level_scales):
        out, _, _ = grid_encode_forward(inputs.detach().numpy(), embeddings.detach().numpy(), geo['offsets'], geo['S'], geo['H'],
                                        geo['gridtype'], geo['align_corners'], False, level_scales)
        ctx.save_for_backward(inputs)
        ctx.geo, ctx.level_scales, ctx.rows = geo, level_scales, embeddings.shape[0]
        L, B, C = out.shape
        return torch.from_numpy(out).permute(1, 0, 2).reshape(B, L * C)

    @staticmethod
    def backward(ctx, grad):
        (inputs,) = ctx.saved_tensors
        geo = ctx.geo
        B = inputs.shape[0]
        g = grad.view(B, geo['L'], geo['C']).permute(1, 0, 2).contiguous().numpy()
        ge, _ = grid_encode_backward(g, inputs.numpy(), geo['offsets'], ctx.rows, geo['S'], geo['H'], geo['gridtype'],
                                     geo['align_corners'], None, ctx.level_scales)
        return None, torch.from_numpy(ge), None, None


class _TruncExp(torch.autograd.Function):
    """external/ngp_activation.py:10-21"""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


PARAM_KEYS = ('encoder.embeddings', 'sigma_net.net.0.weight', 'sigma_net.net.0.bias', 'sigma_net.net.1.weight',
              'sigma_net.net.1.bias', 'sigma_net.net.2.weight', 'sigma_net.net.2.bias')


def make_field_params(seed: int = 0, bound: float = 4.0, emb_range: float = 0.5, hidden: int = 64) -> Dict[str, torch.Tensor]:
    """Deterministic NeRFNetwork parameters keyed by reference state_dict names.  The reference
    initialises embeddings U(-1e-4,1e-4) (grid.py:131-133), a featureless blob; benches and parity
    tests use +-emb_range so that every level contributes."""
    geo = live_geometry(bound)
    rng = np.random.default_rng(seed)
    rows = int(geo['offsets'][-1])
    p = {'encoder.embeddings': torch.from_numpy((rng.random((rows, geo['C']), dtype=np.float32) * 2 - 1) * emb_range)}
    dims = [(hidden, geo['L'] * geo['C']), (hidden, hidden), (4, hidden)]
    for i, (o, k) in enumerate(dims):
        b = 1.0 / math.sqrt(k)
        p[f'sigma_net.net.{i}.weight'] = torch.from_numpy((rng.random((o, k), dtype=np.float32) * 2 - 1) * b)
        p[f'sigma_net.net.{i}.bias'] = torch.from_numpy((rng.random((o,), dtype=np.float32) * 2 - 1) * b)
    return p


class Field:
    """NeRFNetwork restated as a function of a parameter dict (network_grid.py:36-88)."""

    def __init__(self, params: Dict[str, torch.Tensor], bound: float = 4.0, level_scales=None):
        self.p = params
        self.bound = bound
        self.geo = live_geometry(bound)
        self.level_scales = level_scales

    def encode(self, x):
        inputs = (x + self.bound) / (2 * self.bound)  # grid.py:142
        return _GridEncode.apply(inputs.reshape(-1, 3).contiguous(), self.p['encoder.embeddings'], self.geo, self.level_scales)

    def mlp(self, h):
        h = F.relu(F.linear(h, self.p['sigma_net.net.0.weight'], self.p['sigma_net.net.0.bias']))
        h = F.relu(F.linear(h, self.p['sigma_net.net.1.weight'], self.p['sigma_net.net.1.bias']))
        return F.linear(h, self.p['sigma_net.net.2.weight'], self.p['sigma_net.net.2.bias'])

    def common_forward(self, x):
        h = self.mlp(self.encode(x))
        blob = 5 * torch.exp(-(x ** 2).sum(-1) / (2 * 0.2 ** 2))  # network_grid.py:69-75
        sigma = _TruncExp.apply(h[..., 0] + blob)
        albedo = torch.sigmoid(h[..., 1:])
        return sigma, albedo


# ---------------------------------------------------------------------------------------------
# renderer_df.run  (default path)
# ---------------------------------------------------------------------------------------------
def sample_pdf(bins, weights, n_samples, det=False, u=None):
    """renderer_df.py:15-49; ``u`` replaces the torch.rand draw at :31"""
    weights = weights + 1e-5
    pdf = weights / torch.sum(weights, -1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if det:
        u = torch.linspace(0. + 0.5 / n_samples, 1. - 0.5 / n_samples, steps=n_samples)
        u = u.expand(list(cdf.shape[:-1]) + [n_samples])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.max(torch.zeros_like(inds - 1), inds - 1)
    above = torch.min((cdf.shape[-1] - 1) * torch.ones_like(inds), inds)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bins_b, bins_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bins_b + t * (bins_a - bins_b)


def run(field: Field, rays_o, rays_d, *, num_steps=64, upsample_steps=64, min_near=0.1, bg_color=0.0, perturb_noise=None,
        pdf_noise=None, training=True, aabb=None, z_sorted_override=None):
    """renderer_df.py:310-468 for shading='albedo', bg_radius=0.  rays_o/d [N,3] torch fp32.
    perturb_noise [N,num_steps] U(0,1) or None (perturb=False); pdf_noise [N,upsample_steps] U(0,1)
    (ignored when not training: det sampling, :392)."""
    rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    N = rays_o.shape[0]
    b = field.bound
    aabb = torch.tensor([-b, -b, -b, b, b, b], dtype=torch.float32) if aabb is None else aabb
    nears, fars = near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb.numpy(), min_near)
    nears, fars = torch.from_numpy(nears)[:, None], torch.from_numpy(fars)[:, None]
    z_vals = torch.linspace(0.0, 1.0, num_steps).unsqueeze(0).expand((N, num_steps))
    z_vals = nears + (fars - nears) * z_vals
    sample_dist = (fars - nears) / num_steps
    if perturb_noise is not None:
        z_vals = z_vals + (perturb_noise - 0.5) * sample_dist
    xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z_vals.unsqueeze(-1)
    xyzs = torch.min(torch.max(xyzs, aabb[:3]), aabb[3:])
    sigma, albedo = field.common_forward(xyzs.reshape(-1, 3))
    sigma = sigma.view(N, num_steps)
    if upsample_steps > 0:
        with torch.no_grad():
            deltas = z_vals[..., 1:] - z_vals[..., :-1]
            deltas = torch.cat([deltas, sample_dist * torch.ones_like(deltas[..., :1])], dim=-1)
            alphas = 1 - torch.exp(-deltas * sigma)
            alphas_shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
            weights = alphas * torch.cumprod(alphas_shifted, dim=-1)[..., :-1]
            z_vals_mid = (z_vals[..., :-1] + 0.5 * deltas[..., :-1])
            new_z = sample_pdf(z_vals_mid, weights[:, 1:-1], upsample_steps, det=not training, u=pdf_noise).detach()
            new_xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * new_z.unsqueeze(-1)
            new_xyzs = torch.min(torch.max(new_xyzs, aabb[:3]), aabb[3:])
        z_vals = torch.cat([z_vals, new_z], dim=1)
        z_vals, z_index = torch.sort(z_vals, dim=1)
        xyzs = torch.cat([xyzs, new_xyzs], dim=1)
        xyzs = torch.gather(xyzs, dim=1, index=z_index.unsqueeze(-1).expand_as(xyzs))
        if z_sorted_override is not None:
            # the inverse-CDF resampling is ill-conditioned where the pdf sits on its 1e-5 floor; to compare the stages AFTER it
            # (field at the samples, compositing, all gradients) tightly, both sides can be fed the same merged depths
            z_vals = z_sorted_override
            xyzs = rays_o.unsqueeze(-2) + rays_d.unsqueeze(-2) * z_vals.unsqueeze(-1)
            xyzs = torch.min(torch.max(xyzs, aabb[:3]), aabb[3:])
    # The reference evaluates density(new_xyzs) and gathers the coarse+fine sigmas through the sort
    # (:398-412); those sigmas only feed `weights`, and the colour pass (:424) re-evaluates the field
    # at the identical sorted points.  Evaluating once at the sorted points gives the same values
    # and -- because both reference passes share the parameters -- the same parameter gradients.
    T = z_vals.shape[1]
    sig, rgbs = field.common_forward(xyzs.reshape(-1, 3))
    sig, rgbs = sig.view(N, T), rgbs.view(N, T, 3)
    deltas = z_vals[..., 1:] - z_vals[..., :-1]
    deltas = torch.cat([deltas, sample_dist * torch.ones_like(deltas[..., :1])], dim=-1)
    alphas = 1 - torch.exp(-deltas * sig)
    alphas_shifted = torch.cat([torch.ones_like(alphas[..., :1]), 1 - alphas + 1e-15], dim=-1)
    weights = alphas * torch.cumprod(alphas_shifted, dim=-1)[..., :-1]
    weights_sum = weights.sum(dim=-1)
    ori_z = ((z_vals - nears) / (fars - nears)).clamp(0, 1)
    depth = torch.sum(weights * ori_z, dim=-1)
    image = torch.sum(weights.unsqueeze(-1) * rgbs, dim=-2)
    image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
    return dict(image=image, depth=depth, weights_sum=weights_sum, mask=(nears < fars).reshape(-1), z_vals=z_vals,
                nears=nears[:, 0], fars=fars[:, 0])


# ---------------------------------------------------------------------------------------------
# renderer_df.run_cuda / update_extra_state  (cuda_ray mode)
# ---------------------------------------------------------------------------------------------
def run_cuda_train(field: Field, rays_o, rays_d, bitfield, *, cascade=3, grid_size=128, noises=None, dt_gamma=0.0,
                   max_steps=256, T_thresh=1e-4, bg_color=0.0):
    """renderer_df.py:493-507 + :559-584 with force_all_rays (distillation.py:209).  NB run_cuda calls
    near_far_from_aabb without min_near -> 0.2 (renderer_df.py:483)."""
    rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    N = rays_o.shape[0]
    b = field.bound
    aabb = np.array([-b, -b, -b, b, b, b], np.float32)
    nears, fars = near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb, 0.2)
    noises = np.zeros(N, np.float32) if noises is None else noises
    xyzs, dirs, deltas, rays, counter = march_rays_train(rays_o.numpy(), rays_d.numpy(), b, bitfield, cascade, grid_size, nears, fars,
                                                         noises, dt_gamma, max_steps)
    m = int(counter[0])
    m += 128 - m % 128  # raymarching.py:225-229 (align=128)
    xyzs, deltas = xyzs[:m], deltas[:m]
    sigmas, rgbs = field.common_forward(torch.from_numpy(xyzs))

    class _Comp(torch.autograd.Function):
        @staticmethod
        def forward(ctx, s, c):
            ws, depth, image = composite_rays_train_forward(s.detach().numpy(), c.detach().numpy(), deltas, rays, T_thresh)
            ctx.save_for_backward(s, c, torch.from_numpy(ws), torch.from_numpy(image))
            return torch.from_numpy(ws), torch.from_numpy(depth), torch.from_numpy(image)

        @staticmethod
        def backward(ctx, gws, gd, gi):
            s, c, ws, image = ctx.saved_tensors
            gs, gc = composite_rays_train_backward(gws.contiguous().numpy(), gi.contiguous().numpy(), s.detach().numpy(),
                                                   c.detach().numpy(), deltas, rays, ws.numpy(), image.numpy(), T_thresh)
            return torch.from_numpy(gs), torch.from_numpy(gc)

    ws, depth, image = _Comp.apply(sigmas, rgbs)
    image = image + (1 - ws).unsqueeze(-1) * bg_color
    nt, ft = torch.from_numpy(nears), torch.from_numpy(fars)
    depth = torch.clamp(depth - nt, min=0) / (ft - nt)
    return dict(image=image, depth=depth, weights_sum=ws, mask=nt < ft, rays=rays, n_points=int(counter[0]),
                xyzs=xyzs, deltas=deltas)


def run_cuda_eval(field: Field, rays_o, rays_d, bitfield, *, cascade=3, grid_size=128, dt_gamma=0.0, max_steps=256,
                  T_thresh=1e-4, bg_color=0.0):
    """renderer_df.py:521-557 (perturb False)"""
    rays_o, rays_d = rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)
    N = rays_o.shape[0]
    b = field.bound
    aabb = np.array([-b, -b, -b, b, b, b], np.float32)
    nears, fars = near_far_from_aabb(rays_o.numpy(), rays_d.numpy(), aabb, 0.2)
    ws, depth, image = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    rays_alive = np.arange(N, dtype=np.int32)
    rays_t = nears.copy()
    step = 0
    with torch.no_grad():
        while step < max_steps:
            n_alive = rays_alive.shape[0]
            if n_alive <= 0:
                break
            n_step = max(min(N // n_alive, 8), 1)
            xyzs, dirs, deltas = march_rays(n_alive, n_step, rays_alive, rays_t, rays_o.numpy(), rays_d.numpy(), b, bitfield, cascade,
                                            grid_size, nears, fars, np.zeros(n_alive, np.float32), dt_gamma, max_steps, 128)
            sigmas, rgbs = field.common_forward(torch.from_numpy(xyzs))
            composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas.numpy(), rgbs.numpy(), deltas, ws, depth, image, T_thresh)
            rays_alive = np.ascontiguousarray(rays_alive[rays_alive >= 0])
            step += n_step
    wst, nt, ft = torch.from_numpy(ws), torch.from_numpy(nears), torch.from_numpy(fars)
    img = torch.from_numpy(image) + (1 - wst).unsqueeze(-1) * bg_color
    dep = torch.clamp(torch.from_numpy(depth) - nt, min=0) / (ft - nt)
    return dict(image=img, depth=dep, weights_sum=wst, mask=nt < ft)


def update_extra_state(field: Field, density_grid, jitter, *, cascade=3, grid_size=128, decay=0.95, density_thresh=10.0):
    """renderer_df.py:587-640.  density_grid [cascade, G^3] float32 (updated copy returned); jitter
    [cascade, G^3, 3] U(0,1) in (x-major meshgrid) coordinate order replaces torch.rand_like at :618.
    Returns (density_grid, mean_density, bitfield)."""
    G = grid_size
    ar = torch.arange(G, dtype=torch.int32)
    xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing='ij')
    coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
    indices = torch.from_numpy(morton3D(coords.numpy())).long()
    xyzs = 2 * coords.float() / (G - 1) - 1
    tmp = -torch.ones(cascade, G ** 3)
    with torch.no_grad():
        for cas in range(cascade):
            bound = min(2 ** cas, field.bound)
            hgs = bound / G
            cas_xyzs = xyzs * (bound - hgs)
            cas_xyzs = cas_xyzs + (torch.as_tensor(jitter[cas]) * 2 - 1) * hgs
            sig, _ = field.common_forward(cas_xyzs)
            tmp[cas, indices] = sig.reshape(-1)
    grid = torch.as_tensor(density_grid).clone()
    valid = grid >= 0
    grid[valid] = torch.maximum(grid[valid] * decay, tmp[valid])
    mean_density = torch.mean(grid[valid]).item()
    thresh = min(mean_density, density_thresh)
    return grid, mean_density, packbits(grid.numpy(), thresh)


# ---------------------------------------------------------------------------------------------
# synthetic cameras / rays (harness-defined boundary: SURVEY.md §8c, §8d C2)
# ---------------------------------------------------------------------------------------------
def circle_cameras(n_views: int, radius: float = 5.0, elevation_deg: float = 15.0):
    """object-centric fly-around: camera centres on a circle, looking at the origin, +y up"""
    cams = []
    el = math.radians(elevation_deg)
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
    """Pixel-centre grid in NDC from 1-1/W to -1+1/W (utils/render_utils.py:40-47), un-normalised
    directions (plane at depth 1), as pytorch3d's ray sampler produces."""
    c, R = cam
    xs = np.linspace(1 - 1 / W, -1 + 1 / W, W, dtype=np.float32)
    ys = np.linspace(1 - 1 / H, -1 + 1 / H, H, dtype=np.float32)
    yy, xx = np.meshgrid(ys, xs, indexing='ij')
    d_cam = np.stack([xx / focal_ndc, yy / focal_ndc, np.ones_like(xx)], axis=-1).reshape(-1, 3)
    d = d_cam @ R
    o = np.broadcast_to(c, d.shape)
    return np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32)


# ---------------------------------------------------------------------------------------------
# golden vectors
# ---------------------------------------------------------------------------------------------
def golden_grid_inputs(n_random=1024, seed=11):
    """random points plus adversarial ones: 0, 1, out of range, cell boundaries of several levels"""
    rng = np.random.default_rng(seed)
    geo = live_geometry()
    pts = [rng.random((n_random, 3), dtype=np.float32)]
    pts.append(np.array([[0, 0, 0], [1, 1, 1], [0, 1, 0.5], [1.0000001, 0.5, 0.5], [-1e-7, 0.5, 0.5], [0.5, 0.5, 0.5]], np.float32))
    scales = level_scales_host(geo['L'], geo['S'], geo['H'])
    for l in (0, 3, 5, 7, 10, 15):  # x*scale+0.5 exactly integral (when representable) and neighbours
        k = rng.integers(1, int(scales[l]), size=(64, 3))
        base = ((k - 0.5) / scales[l]).astype(np.float32)
        pts += [base, np.nextafter(base, np.float32(1)), np.nextafter(base, np.float32(0))]
    return np.clip(np.concatenate(pts, 0), -1e-6, 1.000001).astype(np.float32)


def write_golden(gold_dir: str):
    geo = live_geometry()
    p = make_field_params(seed=0)
    emb = p['encoder.embeddings'].numpy()
    x = golden_grid_inputs()
    out, dy, rows = grid_encode_forward(x, emb, geo['offsets'], geo['S'], geo['H'], geo['gridtype'], False, True, None, True)
    rng = np.random.default_rng(5)
    g = rng.standard_normal(out.shape, dtype=np.float32)
    ge, gi = grid_encode_backward(g, x, geo['offsets'], emb.shape[0], geo['S'], geo['H'], geo['gridtype'], False, dy, None)
    nz = np.flatnonzero(np.abs(ge).sum(1))[::7]  # a strided sample keeps the fixture small
    np.savez_compressed(os.path.join(gold_dir, 'ngp_grid.npz'), x=x, out=out, rows=rows, grad_seed=5, ge_rows=nz.astype(np.int32),
                        ge_vals=ge[nz], ge_abs_sum=float(np.abs(ge).sum()), gi=gi, level_scales=level_scales_host(geo['L'], geo['S'], geo['H']), offsets=geo['offsets'])
    # hash-type grid on a small geometry (op contract, not live)
    offs, pls = grid_geometry(3, 8, 2, 2.0, 16, 14, None, False)
    emb_h = (rng.random((int(offs[-1]), 2), dtype=np.float32) - 0.5)
    oh, _, rh = grid_encode_forward(x[:1024], emb_h, offs, float(np.log2(pls)), 16, 0, False, False, None, True)
    np.savez_compressed(os.path.join(gold_dir, 'ngp_grid_hash.npz'), x=x[:1024], out=oh, rows=rh, offsets=offs, emb_seed=5,
                        level_scales=level_scales_host(8, float(np.log2(pls)), 16))
    # raymarching utils + cuda_ray pipeline on a 32x32 view
    cam = circle_cameras(8)[1]
    ro, rd = camera_rays(cam, 32, 32)
    aabb = np.array([-4, -4, -4, 4, 4, 4], np.float32)
    nears, fars = near_far_from_aabb(ro, rd, aabb, 0.2)
    field = Field(p)
    grid0 = np.zeros((3, 128 ** 3), np.float32)
    jitter = np.random.default_rng(9).random((3, 128 ** 3, 3), dtype=np.float32)
    grid1, mean_density, bitfield = update_extra_state(field, grid0, jitter)
    noises = np.random.default_rng(10).random(ro.shape[0], dtype=np.float32)
    xyzs, dirs, deltas, rays, counter = march_rays_train(ro, rd, 4.0, bitfield, 3, 128, nears, fars, noises, 0.0, 256)
    res = run_cuda_train(field, torch.from_numpy(ro), torch.from_numpy(rd), bitfield, noises=noises)
    ev = run_cuda_eval(field, torch.from_numpy(ro), torch.from_numpy(rd), bitfield)
    np.savez_compressed(os.path.join(gold_dir, 'ngp_march.npz'), rays_o=ro, rays_d=rd, nears=nears, fars=fars,
                        bitfield=bitfield, mean_density=mean_density, noises=noises, rays=rays, counter=counter,
                        n_points=int(counter[0]), image=res['image'].detach().numpy(), weights_sum=res['weights_sum'].detach().numpy(),
                        depth=res['depth'].detach().numpy(), eval_image=ev['image'].numpy(), eval_ws=ev['weights_sum'].numpy())
    # default run() path on the same 32x32 view, with gradients of a fixed loss
    N = ro.shape[0]
    pn = np.random.default_rng(12).random((N, 64), dtype=np.float32)
    un = np.random.default_rng(13).random((N, 64), dtype=np.float32)
    params = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    f2 = Field(params)
    r = run(f2, torch.from_numpy(ro), torch.from_numpy(rd), perturb_noise=torch.from_numpy(pn), pdf_noise=torch.from_numpy(un))
    tgt = torch.from_numpy(np.random.default_rng(14).random((N, 3), dtype=np.float32))
    loss = ((r['image'] - tgt) ** 2).mean() + 0.1 * r['weights_sum'].mean()
    loss.backward()
    gemb = params['encoder.embeddings'].grad.numpy()
    nzr = np.flatnonzero(np.abs(gemb).sum(1))[::61]  # strided sample of the touched rows
    np.savez_compressed(os.path.join(gold_dir, 'ngp_run.npz'), rays_o=ro, rays_d=rd, perturb_seed=12, pdf_seed=13, target_seed=14,
                        gemb_abs_sum=float(np.abs(gemb).sum()),
                        image=r['image'].detach().numpy(), weights_sum=r['weights_sum'].detach().numpy(), depth=r['depth'].detach().numpy(),
                        z_vals=r['z_vals'].numpy()[::8], loss=float(loss.detach()), gemb_rows=nzr.astype(np.int32), gemb_vals=gemb[nzr],
                        **{'g_' + k: params[k].grad.numpy() for k in PARAM_KEYS[1:]})
    print('[ngp golden] grid', out.shape, 'march points', int(counter[0]), 'run loss', float(loss))
