"""NeRFRenderer -- host-side mirror of external/nerf/renderer_df.py (reference), the volume renderer the
distillation loop drives (sparsefusion/distillation.py:209,282,380).

Same class / method names and keyword arguments (``render``, ``render_batched``, ``run``, ``run_cuda``,
``update_extra_state``, ``reset_extra_state``), same result dictionaries, same registered buffers
(``aabb_train``, ``aabb_infer``, and ``density_grid`` / ``density_bitfield`` / ``step_counter`` with cuda_ray), so
``ngp_network.render(rays_o, rays_d, staged=False, perturb=True, bg_color=0, ..., **vars(opt))`` works unchanged.
``export_mesh`` (offline asset export through mcubes/xatlas) is out of scope.

``run`` is the default path (opt.cuda_ray = False, distillation.py:505).  The reference spells it as ~150 eager
launches and evaluates the field three times (coarse, fine, colour); here it is seven kernels and every point is
evaluated once per role (coarse density for the sampling pdf, then sigma+rgb at the 128 sorted samples), with an
analytic backward -- gradients equal the reference's because its fine/colour passes see the same points and
parameters (DESIGN.md spells out the argument).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import _lib as lib
from . import raymarching


class _RunRender(torch.autograd.Function):
    """renderer_df.run (:310-468) for shading='albedo', bg_radius == 0: returns (image [N,3], depth [N], weights_sum [N], nears, fars)"""

    @staticmethod
    def forward(ctx, net, rays_o, rays_d, aabb, min_near, bg_color, perturb_noise, pdf_u, emb, w0, b0, w1, b1, w2, b2):
        N = rays_o.shape[0]
        dev = rays_o.device
        enc = net.encoder
        S, H, bound = float(math.log2(enc.per_level_scale)), int(enc.base_resolution), float(net.bound)
        st = lib.stream()
        f = lib.fptr
        nears, fars = torch.empty(N, device=dev), torch.empty(N, device=dev)
        zc = torch.empty(N, 64, device=dev)
        lin = torch.linspace(0.0, 1.0, 64, device=dev)
        lib.call('sfb_ray_coarse_z', f(rays_o), f(rays_d), f(aabb), float(min_near), f(lin), f(perturb_noise), N, 64, f(nears), f(fars), f(zc), st)
        # the field is evaluated ONCE per sample: coarse pass (sigma + rgb at the 64 stratified depths), importance pass (the 64 new depths),
        # then a gather into sorted order -- the values of the coarse samples are the same numbers a second evaluation would produce
        sig_c, rgb_c = torch.empty(N, 64, device=dev), torch.empty(N, 64, 3, device=dev)
        args = (f(emb), lib.iptr(enc.offsets), S, H, bound, f(w0), f(b0), f(w1), f(b1), f(w2), f(b2))
        lib.call('sfb_ngp_field_forward', None, f(rays_o), f(rays_d), f(zc), 64, N * 64, *args, f(sig_c), f(rgb_c), st)
        zs, zn = torch.empty(N, 128, device=dev), torch.empty(N, 64, device=dev)
        src_of = torch.empty(N, 128, dtype=torch.uint8, device=dev)
        lib.call('sfb_ray_resample_ex', f(zc), f(sig_c), f(nears), f(fars), f(pdf_u), 0, N, 64, 64, f(zs), f(zn), src_of.data_ptr(), st)
        sig_n, rgb_n = torch.empty(N, 64, device=dev), torch.empty(N, 64, 3, device=dev)
        lib.call('sfb_ngp_field_forward', None, f(rays_o), f(rays_d), f(zn), 64, N * 64, *args, f(sig_n), f(rgb_n), st)
        sigma, rgb = torch.empty(N, 128, device=dev), torch.empty(N, 128, 3, device=dev)
        lib.call('sfb_ray_gather_sorted', src_of.data_ptr(), f(sig_c), f(rgb_c), f(sig_n), f(rgb_n), N, 128, f(sigma), f(rgb), st)
        image, depth, ws = torch.empty(N, 3, device=dev), torch.empty(N, device=dev), torch.empty(N, device=dev)
        lib.call('sfb_ray_composite_forward', f(zs), f(sigma), f(rgb), f(nears), f(fars), float(bg_color), N, 128, f(image), f(depth), f(ws), st)
        ctx.save_for_backward(rays_o, rays_d, zs, sigma, rgb, nears, fars, emb, w0, b0, w1, b1, w2, b2)
        ctx.net, ctx.bg = net, float(bg_color)
        ctx.mark_non_differentiable(nears, fars)
        return image, depth, ws, nears, fars

    @staticmethod
    def backward(ctx, g_image, g_depth, g_ws, _gn, _gf):
        rays_o, rays_d, zs, sigma, rgb, nears, fars, emb, w0, b0, w1, b1, w2, b2 = ctx.saved_tensors
        net = ctx.net
        enc = net.encoder
        N = rays_o.shape[0]
        dev = rays_o.device
        S, H, bound = float(math.log2(enc.per_level_scale)), int(enc.base_resolution), float(net.bound)
        st = lib.stream()
        f = lib.fptr
        g_image = g_image.contiguous() if g_image is not None else torch.zeros(N, 3, device=dev)
        g_sigma, g_rgb = torch.empty(N, 128, device=dev), torch.empty(N, 128, 3, device=dev)
        g_ws = None if g_ws is None else g_ws.contiguous()
        g_depth = None if g_depth is None else g_depth.contiguous()
        lib.call('sfb_ray_composite_backward', f(zs), f(sigma), f(rgb), f(nears), f(fars), ctx.bg, N, 128, f(g_image),
                 None if g_ws is None else f(g_ws), None if g_depth is None else f(g_depth), f(g_sigma), f(g_rgb), st)
        B = N * 128
        tape = torch.empty(lib.load().sfb_ngp_field_tape_floats(B), device=dev)
        g_emb = torch.zeros_like(emb)
        gw0, gb0, gw1, gb1, gw2, gb2 = (torch.zeros_like(t) for t in (w0, b0, w1, b1, w2, b2))
        lib.call('sfb_ngp_field_backward', None, f(rays_o), f(rays_d), f(zs), 128, B, f(emb), lib.iptr(enc.offsets), S, H, bound, f(w0), f(b0), f(w1),
                 f(b1), f(w2), f(b2), f(g_sigma), f(g_rgb), f(g_emb), f(gw0), f(gb0), f(gw1), f(gb1), f(gw2), f(gb2), f(tape), st)
        return (None, None, None, None, None, None, None, None, g_emb, gw0, gb0, gw1, gb1, gw2, gb2)


class NeRFRenderer(nn.Module):
    MAX_RAYS_PER_LAUNCH = 65536      # rays per fused-render autograd node (x 128 samples x 1.4 KB of backward tape = 12 GB transient)

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.bound = opt.bound
        self.cascade = 1 + math.ceil(math.log2(opt.bound))
        self.grid_size = 128
        self.cuda_ray = opt.cuda_ray
        self.min_near = opt.min_near
        self.density_thresh = opt.density_thresh
        self.bg_radius = opt.bg_radius
        aabb_train = torch.FloatTensor([-opt.bound, -opt.bound, -opt.bound, opt.bound, opt.bound, opt.bound])
        self.register_buffer('aabb_train', aabb_train)
        self.register_buffer('aabb_infer', aabb_train.clone())
        if self.cuda_ray:
            self.register_buffer('density_grid', torch.zeros([self.cascade, self.grid_size ** 3]))
            self.register_buffer('density_bitfield', torch.zeros(self.cascade * self.grid_size ** 3 // 8, dtype=torch.uint8))
            self.mean_density = 0
            self.iter_density = 0
            self.register_buffer('step_counter', torch.zeros(16, 2, dtype=torch.int32))
            self.mean_count = 0
            self.local_step = 0

    def forward(self, x, d):
        raise NotImplementedError()

    def density(self, x):
        raise NotImplementedError()

    def reset_extra_state(self):
        if not self.cuda_ray:
            return
        self.density_grid.zero_()
        self.mean_density = 0
        self.iter_density = 0
        self.step_counter.zero_()
        self.mean_count = 0
        self.local_step = 0

    def export_mesh(self, *a, **k):
        raise NotImplementedError('export_mesh (mcubes / xatlas asset export) is outside the distillation hot path')

    # ------------------------------------------------------------------------------------------ default path
    def run(self, rays_o, rays_d, num_steps=128, upsample_steps=128, light_d=None, ambient_ratio=1.0, shading='albedo', bg_color=None,
            perturb=False, fixed_light=False, perturb_noise=None, pdf_noise=None, **kwargs):
        """rays_o, rays_d: [B, N, 3] (B == 1) -> image [B,N,3], depth [B,N], weights_sum [N], mask [B,N].
        ``perturb_noise`` [N,num_steps] / ``pdf_noise`` [N,upsample_steps] override the torch.rand draws (parity tests)."""
        if shading != 'albedo' or self.bg_radius > 0:
            raise NotImplementedError("the fused renderer covers SparseFusion's configuration: shading='albedo', bg_radius=0 "
                                      '(sparsefusion/distillation.py:209,512)')
        if num_steps != 64 or upsample_steps != 64:
            raise NotImplementedError('the fused renderer is built for num_steps = upsample_steps = 64 (get_default_torch_ngp_opt)')
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3).float()
        rays_d = rays_d.contiguous().view(-1, 3).float()
        if not rays_o.is_cuda:
            raise RuntimeError('NeRFRenderer.run needs CUDA tensors (the reference has no CPU path either)')
        N = rays_o.shape[0]
        dev = rays_o.device
        aabb = self.aabb_train if self.training else self.aabb_infer
        if bg_color is None:
            bg_color = 1
        if not isinstance(bg_color, (int, float)):
            raise NotImplementedError('per-ray bg_color tensors are not used by the distillation loop (bg_color=0)')
        with torch.cuda.device(dev):
            if perturb and perturb_noise is None:
                perturb_noise = torch.rand(N, num_steps, device=dev)                          # renderer_df.py:363
            if not perturb:
                perturb_noise = None
            if self.training:
                pdf_u = pdf_noise if pdf_noise is not None else torch.rand(N, upsample_steps, device=dev)   # renderer_df.py:31
            else:                                                                              # det = not self.training (:392)
                pdf_u = torch.linspace(0. + 0.5 / upsample_steps, 1. - 0.5 / upsample_steps, steps=upsample_steps, device=dev).expand(N, upsample_steps)
            pdf_u = pdf_u.contiguous()
            p = self._field_params()
            pn = None if perturb_noise is None else perturb_noise.contiguous()
            if N <= self.MAX_RAYS_PER_LAUNCH:
                image, depth, ws, nears, fars = _RunRender.apply(self, rays_o, rays_d, aabb, self.min_near, bg_color, pn, pdf_u, *p)
            else:
                # large images (512x512 rays and up): the backward's activation tapes are 1.4 KB per sample point, so rays go through the fused
                # kernels in slices -- each slice is its own autograd node, whose tape exists only while that slice back-propagates
                parts = [_RunRender.apply(self, rays_o[i:i + self.MAX_RAYS_PER_LAUNCH], rays_d[i:i + self.MAX_RAYS_PER_LAUNCH], aabb, self.min_near, bg_color,
                                          None if pn is None else pn[i:i + self.MAX_RAYS_PER_LAUNCH], pdf_u[i:i + self.MAX_RAYS_PER_LAUNCH], *p)
                         for i in range(0, N, self.MAX_RAYS_PER_LAUNCH)]
                image, depth, ws, nears, fars = (torch.cat([q[j] for q in parts]) for j in range(5))
        return {'image': image.view(*prefix, 3), 'depth': depth.view(*prefix), 'weights_sum': ws, 'mask': (nears < fars).reshape(*prefix)}

    # ------------------------------------------------------------------------------------------ cuda_ray path
    def run_cuda(self, rays_o, rays_d, dt_gamma=0, light_d=None, ambient_ratio=1.0, shading='albedo', bg_color=None, perturb=False,
                 force_all_rays=False, max_steps=1024, T_thresh=1e-4, **kwargs):
        prefix = rays_o.shape[:-1]
        rays_o = rays_o.contiguous().view(-1, 3)
        rays_d = rays_d.contiguous().view(-1, 3)
        N = rays_o.shape[0]
        device = rays_o.device
        # NB: like the reference (renderer_df.py:483) min_near is NOT forwarded here -> the wrapper default 0.2
        nears, fars = raymarching.near_far_from_aabb(rays_o, rays_d, self.aabb_train if self.training else self.aabb_infer)
        results = {}
        if self.training:
            counter = self.step_counter[self.local_step % 16]
            counter.zero_()
            self.local_step += 1
            xyzs, dirs, deltas, rays = raymarching.march_rays_train(rays_o, rays_d, self.bound, self.density_bitfield, self.cascade, self.grid_size,
                                                                    nears, fars, counter, self.mean_count, perturb, 128, force_all_rays, dt_gamma,
                                                                    max_steps)
            sigmas, rgbs, _ = self(xyzs, dirs, light_d, ratio=ambient_ratio, shading=shading)
            weights_sum, depth, image = raymarching.composite_rays_train(sigmas, rgbs, deltas, rays, T_thresh)
        else:
            weights_sum = torch.zeros(N, dtype=torch.float32, device=device)
            depth = torch.zeros(N, dtype=torch.float32, device=device)
            image = torch.zeros(N, 3, dtype=torch.float32, device=device)
            n_alive = N
            rays_alive = torch.arange(n_alive, dtype=torch.int32, device=device)
            rays_t = nears.clone()
            step = 0
            while step < max_steps:
                n_alive = rays_alive.shape[0]
                if n_alive <= 0:
                    break
                n_step = max(min(N // n_alive, 8), 1)
                xyzs, dirs, deltas = raymarching.march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, self.bound, self.density_bitfield,
                                                            self.cascade, self.grid_size, nears, fars, 128, perturb if step == 0 else False,
                                                            dt_gamma, max_steps)
                sigmas, rgbs, _ = self(xyzs, dirs, light_d, ratio=ambient_ratio, shading=shading)
                raymarching.composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh)
                rays_alive = rays_alive[rays_alive >= 0]
                step += n_step
        if self.bg_radius > 0:
            raise NotImplementedError('bg_radius > 0 is not used by SparseFusion (distillation.py:512)')
        if bg_color is None:
            bg_color = 1
        image = image + (1 - weights_sum).unsqueeze(-1) * bg_color
        image = image.view(*prefix, 3)
        depth = torch.clamp(depth - nears, min=0) / (fars - nears)
        results['image'] = image
        results['depth'] = depth.view(*prefix)
        results['weights_sum'] = weights_sum.reshape(*prefix)
        results['mask'] = (nears < fars).reshape(*prefix)
        return results

    @torch.no_grad()
    def update_extra_state(self, decay=0.95, S=128, jitter=None):
        """renderer_df.py:587-640: refresh the cascaded density grid (EMA max) and its bitfield.  ``jitter``
        [cascade, G^3, 3] in U(0,1) overrides torch.rand_like (parity tests)."""
        if not self.cuda_ray:
            return
        G = self.grid_size
        dev = self.density_bitfield.device
        tmp_grid = -torch.ones_like(self.density_grid)
        ar = torch.arange(G, dtype=torch.int32, device=dev)
        xx, yy, zz = torch.meshgrid(ar, ar, ar, indexing='ij')
        coords = torch.stack([xx.reshape(-1), yy.reshape(-1), zz.reshape(-1)], dim=-1)
        indices = raymarching.morton3D(coords).long()
        xyzs = 2 * coords.float() / (G - 1) - 1
        for cas in range(self.cascade):
            bound = min(2 ** cas, self.bound)
            half_grid_size = bound / G
            cas_xyzs = xyzs * (bound - half_grid_size)
            noise = torch.rand_like(cas_xyzs) if jitter is None else jitter[cas].to(dev)
            cas_xyzs = cas_xyzs + (noise * 2 - 1) * half_grid_size
            sigmas = self.density(cas_xyzs)['sigma'].reshape(-1).detach()
            tmp_grid[cas, indices] = sigmas
        valid_mask = self.density_grid >= 0
        self.density_grid[valid_mask] = torch.maximum(self.density_grid[valid_mask] * decay, tmp_grid[valid_mask])
        self.mean_density = torch.mean(self.density_grid[valid_mask]).item()
        self.iter_density += 1
        density_thresh = min(self.mean_density, self.density_thresh)
        self.density_bitfield = raymarching.packbits(self.density_grid, density_thresh, self.density_bitfield)
        total_step = min(16, self.local_step)
        if total_step > 0:
            self.mean_count = int(self.step_counter[:total_step, 0].sum().item() / total_step)
        self.local_step = 0

    # ------------------------------------------------------------------------------------------ dispatch
    def render(self, rays_o, rays_d, staged=False, max_ray_batch=4096, **kwargs):
        _run = self.run_cuda if self.cuda_ray else self.run
        B, N = rays_o.shape[:2]
        device = rays_o.device
        if staged and not self.cuda_ray:
            depth = torch.empty((B, N), device=device)
            image = torch.empty((B, N, 3), device=device)
            weights_sum = torch.empty((B, N), device=device)
            for b in range(B):
                head = 0
                while head < N:
                    tail = min(head + max_ray_batch, N)
                    r = _run(rays_o[b:b + 1, head:tail], rays_d[b:b + 1, head:tail], **kwargs)
                    depth[b:b + 1, head:tail] = r['depth']
                    weights_sum[b:b + 1, head:tail] = r['weights_sum']
                    image[b:b + 1, head:tail] = r['image']
                    head += max_ray_batch
            return {'depth': depth, 'image': image, 'weights_sum': weights_sum}
        return _run(rays_o, rays_d, **kwargs)

    def render_batched(self, rays_o, rays_d, batched=False, max_ray_batch=128 * 128, **kwargs):
        with torch.no_grad():
            return self.render(rays_o, rays_d, staged=batched, max_ray_batch=max_ray_batch, **kwargs)
