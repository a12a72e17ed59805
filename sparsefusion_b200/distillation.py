"""Score-distillation inner loop -- host-side mirror of the loop body of sparsefusion/distillation.py:174-352 (reference).

The reference's ``distillation_loop`` interleaves the hot loop with dataset / pytorch3d / EFT / matplotlib plumbing.
This module keeps the LOOP BODY -- photometric sub-step A (:185-256) and fusion sub-step B (:259-352) -- as
``Distiller.step(itr)`` over a ``SceneCache`` of device tensors, with the boundary the survey fixes:
rays arrive as ``rays_o / rays_d`` tensors (the reference gets them from pytorch3d's GridRaysampler, :201-204,:274-277)
and the per-view EFT features / images arrive cached (:95-125).  Loss formulas, weights, schedules and the order of
RNG draws follow the reference line by line (citations inline).  ``INTEGRATION.md`` shows how the reference's
``distillation_loop`` calls this.

What runs where: both NGP renders (fused sm_100a kernels, analytic backward), the PLMS sampler's UNet evaluations
(tcgen05 engine, one CUDA graph replay per evaluation), the SD VAE encode / decode (``ldm_autoencoder.py``, same engine), the
image-space losses with their gradients (``image_glue.py``), the LPIPS-VGG term when a ``percep`` module is given (``lpips_vgg.py``;
:312-314), fused Adam.

Two step semantics:

* ``Distiller.step`` with ``views_per_step=None`` on one rank is EXACTLY the reference's iteration: sub-step A, Adam, sub-step B on ONE
  target view, Adam (distillation.py:244-247, :345-352).
* ``views_per_step=V`` (BASELINE configs[3]) is the view-batched extension: the photometric sub-step and V target views are
  evaluated against the SAME parameters, gradient = grad(A) + mean_v grad(B_v), ONE Adam step -- mathematically one optimiser step on a
  V-view minibatch, not V sequential reference steps.  The V views are entries 1..V of the step's permutation (the reference takes entry 1);
  rank r distils views ``views[r::R]`` as ONE batch through the VAE and the PLMS sampler (UNet batch = V/R, one shared ``max_thres`` per
  step, as many renders as views), the photometric sub-step is computed by one rank (``itr % R``), and the flat 7.46 MB gradient is summed
  with ONE all-reduce per step.  Every random draw of a view (render jitter, inverse-CDF samples, PLMS noise) comes from a generator
  keyed by (seed, iteration, view), so the step's result does not depend on R (up to fp32 summation order).

``fusion_substep`` / ``photometric_substep`` with world_size > 1 remain as the one-view-per-rank form (rank r takes entry 1 + r of the
permutation, all-reduce, Adam with grad_scale 1/R) that ``tests/test_multirank_gpu.py`` pins against the single-rank minibatch.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib as lib
from . import image_glue as glue
from .plms import PLMSSampler, _log_snr as plms_log_snr, _sigmoid as plms_sigmoid


def normalize(x):  # utils/common_utils.py:9-13
    return torch.clip(x * 2 - 1.0, -1.0, 1.0)


def unnormalize(x):  # utils/common_utils.py:15-19
    return torch.clip((x + 1.0) / 2.0, 0.0, 1.0)


def huber(x, y, scaling=0.1):  # utils/common_utils.py:183-190
    diff_sq = (x - y) ** 2
    return ((1 + diff_sq / (scaling ** 2)).clamp(1e-4).sqrt() - 1) * float(scaling)


def shard_target_view(perm: torch.Tensor, rank: int) -> int:
    """which target view a rank distils this step: entry (1 + rank) of the step's permutation, wrapping (the reference,
    single process, takes entry 1: distillation.py:264).  Ranks < n_targets therefore get distinct views."""
    return int(perm[(1 + rank) % perm.numel()])


def step_views(perm: torch.Tensor, views_per_step: int) -> list:
    """the V target views of a minibatch step: entries 1..V of the step's permutation, wrapping (V = 1 is the reference's choice, :264)"""
    n = perm.numel()
    return [int(perm[(1 + j) % n]) for j in range(views_per_step)]


def shard_views(views: list, rank: int, world_size: int) -> list:
    """rank r's share of a step's views: views[r::R] ({v : v mod R = r} over the step's list)"""
    return list(views[rank::world_size])


def view_seed(seed: int, itr: int, view: int, stream: int) -> int:
    """seed of the generator that serves ONE (iteration, view, stream) -- stream 0: photometric render, 1: fusion render, 2: PLMS noise.
    Keyed draws make a minibatch step independent of how its views are spread over ranks."""
    x = (seed * 0x9E3779B97F4A7C15 + itr * 0xBF58476D1CE4E5B9 + view * 0x94D049BB133111EB + stream * 0xD6E8FEB86659FD93) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 31
    return x & 0x7FFFFFFFFFFFFFFF


class KeyedNoise:
    """``noise_fn`` of the PLMS sampler for a batch of views: row i of every draw comes from view i's own generator (one randn of ``chunk`` draws
    per view per sampling run, sliced by the sampler's calls), so a view's noise does not depend on which other views share its batch"""

    def __init__(self, seeds, device, chunk: int = 128):
        self.gens = []
        for sd in seeds:
            g = torch.Generator(device=device)
            g.manual_seed(int(sd))
            self.gens.append(g)
        self.device, self.chunk, self.buf, self.pos = device, chunk, None, 0

    def __call__(self, like: torch.Tensor) -> torch.Tensor:
        assert like.shape[0] == len(self.gens), 'KeyedNoise: batch rows must match the views it was keyed for'
        if self.buf is None or self.pos >= self.chunk:
            self.buf = torch.stack([torch.randn(self.chunk, *like.shape[1:], generator=g, device=self.device, dtype=like.dtype) for g in self.gens], dim=1)
            self.pos = 0
        out = self.buf[self.pos]
        self.pos += 1
        return out


@dataclass
class SceneCache:
    """what the reference holds per scene before the loop starts (distillation.py:65-125), as device tensors"""
    input_rgb: torch.Tensor        # [Vi,3,256,256]  scene_rgb[input_idx]
    input_mask: torch.Tensor       # [Vi,1,256,256]  scene_mask[input_idx]
    input_rays_o: torch.Tensor     # [Vi,N,3]  sampler_feat(camera_vox) origins, N = 128*128 (:201-204)
    input_rays_d: torch.Tensor     # [Vi,N,3]
    target_features: torch.Tensor  # [Vt,256,32,32]  eft_feature_cache[ci]['features'] (:116,:124)
    target_eft_image: torch.Tensor  # [Vt,3,256,256] eft_feature_cache[ci]['eft_image'] (:118-119)
    target_rays_o: torch.Tensor    # [Vt,N,3]
    target_rays_d: torch.Tensor    # [Vt,N,3]

    def to(self, device, non_blocking=False):
        return SceneCache(*[getattr(self, f).to(device, non_blocking=non_blocking) for f in self.__dataclass_fields__])

    def pin(self):
        return SceneCache(*[getattr(self, f).pin_memory() for f in self.__dataclass_fields__])


class FlatAdam:
    """torch.optim.Adam(ngp.get_params(lr)) + StepLR(step_size, gamma) (distillation.py:165-166) over ONE flat buffer:
    parameters, gradients and both moments are contiguous so that the gradient all-reduce is a single collective and
    the update a single fused kernel per parameter group (encoder lr*10, MLP lr: network_grid.py:223-234)."""

    def __init__(self, net, lr=5e-4, betas=(0.9, 0.999), eps=1e-8, step_size=3000, gamma=0.2):
        groups = [dict(g, params=list(g['params'])) for g in net.get_params(lr)]   # get_params hands out generators
        params = [p for g in groups for p in g['params']]
        dev = params[0].device
        sizes = [p.numel() for p in params]
        self.flat = torch.empty(sum(sizes), device=dev)
        self.grad = torch.zeros_like(self.flat)
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.groups = []
        off = 0
        for g in groups:
            start = off
            for p in g['params']:
                n = p.numel()
                self.flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat[off:off + n].view_as(p)          # parameters become views of the flat buffer
                p.grad = self.grad[off:off + n].view_as(p)          # autograd accumulates in place into the flat gradient
                off += n
            self.groups.append((start, off, g['lr']))
        self.params = params
        self.betas, self.eps, self.step_size, self.gamma = betas, eps, step_size, gamma
        self.t = 0          # optimizer.step() count
        self.sched = 0      # scheduler.step() count

    def zero_grad(self):
        self.grad.zero_()  # parameters' .grad are views of this buffer; never set them to None

    def sync_grads(self, world_size: int = 1, process_group=None) -> float:
        """N>1: ONE sum all-reduce of the flat gradient buffer (NCCL over NVLink on the GPU box, gloo in the CPU tests);
        returns the grad_scale (1/world_size) the following step() must use so that the update is the mean over ranks"""
        if world_size > 1:
            torch.distributed.all_reduce(self.grad, op=torch.distributed.ReduceOp.SUM, group=process_group)
        return 1.0 / world_size

    def step(self, grad_scale: float = 1.0):
        if not self.flat.is_cuda:
            raise RuntimeError('FlatAdam.step: the fused Adam kernel needs CUDA parameters (there is no CPU path)')
        self.t += 1
        lr_mult = self.gamma ** (self.sched // self.step_size)
        for start, end, lr in self.groups:
            lib.call('sfb_adam_step', self.flat[start:end].data_ptr(), self.grad[start:end].data_ptr(), self.m[start:end].data_ptr(),
                     self.v[start:end].data_ptr(), end - start, lr * lr_mult, self.betas[0], self.betas[1], self.eps, self.t, grad_scale, lib.stream())

    def scheduler_step(self):
        self.sched += 1


class Distiller:
    def __init__(self, ngp, vae, vldm, opt, cache: SceneCache, *, z_scale_factor=0.18215, plms_steps=50, start_fusion_step=1000,
                 lambda_color=1.0, lambda_sil=1.0, lambda_opacity=1e-3, seed=0, rank=0, world_size=1, process_group=None,
                 use_cuda_graph=True, fused_glue=True, percep=None, lambda_percep=0.1, start_percep_step=1000, views_per_step=None, max_batch=32):
        self.ngp, self.vae, self.vldm, self.opt, self.cache = ngp, vae, vldm, opt, cache
        self.z_scale_factor = z_scale_factor
        self.start_fusion_step = start_fusion_step
        self.lambda_color, self.lambda_sil, self.lambda_opacity = lambda_color, lambda_sil, lambda_opacity
        self.rank, self.world_size, self.pg = rank, world_size, process_group
        # None: the reference's two-update iteration on one view (one rank) / one view per rank (N ranks, legacy form); V: view-batched minibatch step
        self.views_per_step = views_per_step
        self.max_batch = max_batch     # views per VAE / PLMS batch on one rank (activations of the VAE decoder: 0.27 GB per tensor per 8 views at 256^2)
        self.seed = seed
        # perceptual term (distillation.py:161, :176-178, :312-314): a PerceptualLoss module (lpips_vgg.py) or None; lambda switches on at start_percep_step
        self.percep, self.lambda_percep, self.start_percep_step = percep, lambda_percep, start_percep_step
        self.fused_glue = fused_glue   # image-space losses + their gradients as fused kernels (image_glue.py) instead of ~50 eager launches + autograd
        self.sampler = PLMSSampler(vldm, plms_steps, use_cuda_graph=use_cuda_graph)          # distillation.py:160
        self.optimizer = FlatAdam(ngp, lr=5e-4)                                              # :165-166
        # the reference draws from torch's global CPU generator (:185,:263,:303); a private generator with the same draw
        # order keeps every rank's view permutation identical
        self.gen = torch.Generator().manual_seed(seed)
        self.render_kw = dict(staged=False, perturb=True, bg_color=0, ambient_ratio=1.0, shading='albedo', force_all_rays=True, **vars(opt))
        self.render_noise = None   # parity tests: callable('A'|'B') -> (perturb_noise [N,64], pdf_noise [N,64]) replacing torch.rand
        self.render_noise_keyed = None   # parity tests (minibatch step): callable('A'|'B', view) -> the same pair for that view
        self.plms_noise = None     # parity tests (minibatch step): callable(views) -> noise_fn for that batch, replacing KeyedNoise
        self.pred_img_hook = None  # parity tests (minibatch step): callable(views, pred_img) -> pred_img, e.g. to inject a fixed denoised target
        self.last = {}

    # ------------------------------------------------------------------------------------------------
    def _render_raw(self, rays_o, rays_d, which='A'):
        """(image [N,3], weights_sum [N], hw) with the autograd graph of the render attached"""
        kw = self.render_kw
        if self.render_noise is not None:
            pn, un = self.render_noise(which)
            kw = dict(kw, perturb_noise=pn, pdf_noise=un)
        out = self.ngp.render(rays_o[None], rays_d[None], **kw)
        n = rays_o.shape[0]
        return out['image'].reshape(n, 3), out['weights_sum'].reshape(n), int(round(n ** 0.5))

    def _render(self, rays_o, rays_d, which='A'):
        kw = self.render_kw
        if self.render_noise is not None:
            pn, un = self.render_noise(which)
            kw = dict(kw, perturb_noise=pn, pdf_noise=un)
        out = self.ngp.render(rays_o[None], rays_d[None], **kw)
        hw = int(round(rays_o.shape[0] ** 0.5))
        image = out['image'].reshape(1, hw, hw, 3).permute(0, 3, 1, 2).contiguous()           # :210
        sil = out['weights_sum'].reshape(1, hw, hw, 1).permute(0, 3, 1, 2).contiguous()       # :211
        return image, sil

    def photometric_substep(self, itr: int):
        """distillation.py:185-247"""
        c = self.cache
        n_in = c.input_rgb.shape[0]
        idx = int(torch.randperm(n_in, generator=self.gen)[0])                               # :185-186
        self.ngp.train()
        if self.opt.cuda_ray and itr % 16 == 0:                                               # :181-182
            self.ngp.update_extra_state()
        if self.fused_glue:
            img, ws, hw = self._render_raw(c.input_rays_o[idx], c.input_rays_d[idx], 'A')
            loss, g_img, g_ws = glue.photometric_loss(img.detach(), ws.detach(), c.input_rgb[idx], c.input_mask[idx], hw, hw, int(self.opt.hw_scale),
                                                      self.lambda_color, self.lambda_sil, self.lambda_opacity)        # :216-234 value + gradient
            self.optimizer.zero_grad()                                                        # :244
            torch.autograd.backward((img, ws), (g_img, g_ws))
            self.optimizer.step(grad_scale=self.optimizer.sync_grads(self.world_size, self.pg))
            self.optimizer.scheduler_step()                                                   # :247
            self.last['photo_loss'] = loss
            return loss
        image, sil = self._render(c.input_rays_o[idx], c.input_rays_d[idx], 'A')
        scale = 1.0 / self.opt.hw_scale
        batch_rgb = F.interpolate(c.input_rgb[idx:idx + 1], scale_factor=scale)               # :216 (nearest)
        batch_mask = F.interpolate(c.input_mask[idx:idx + 1], scale_factor=scale)             # :217
        color_err = huber(image, batch_rgb).abs().mean()                                      # :218
        sil_err = huber(sil, batch_mask).abs().mean()                                         # :222
        loss = self.lambda_color * color_err + self.lambda_sil * sil_err                      # :227
        opacity = torch.sqrt(sil ** 2 + .01).mean()                                           # :233
        loss = loss + self.lambda_opacity * opacity                                           # :234
        self.optimizer.zero_grad()                                                            # :244
        loss.backward()
        # replicated sub-step: for N>1 the reduce only removes atomic-order drift so that ranks stay bit-identical
        self.optimizer.step(grad_scale=self.optimizer.sync_grads(self.world_size, self.pg))
        self.optimizer.scheduler_step()                                                       # :247
        self.last['photo_loss'] = loss.detach()
        return loss.detach()

    def fusion_substep(self, itr: int, max_thres: Optional[float] = None):
        """distillation.py:259-352; rank r takes the (1 + r)-th entry of the step's permutation (the reference takes entry 1, :264)"""
        c = self.cache
        self.optimizer.zero_grad()                                                            # :261
        n_t = c.target_features.shape[0]
        perm = torch.randperm(n_t, generator=self.gen)                                        # :263
        vi = shard_target_view(perm, self.rank)                                               # :264
        u = torch.rand(1, generator=self.gen)                                                 # :303 (drawn every step to keep the stream aligned)
        feats = c.target_features[vi:vi + 1]
        if self.fused_glue and int(self.opt.hw_scale) == 2:
            return self._fusion_substep_fused(itr, max_thres, vi, u, feats)
        image, sil = self._render(c.target_rays_o[vi], c.target_rays_d[vi], 'B')
        image = F.interpolate(image, scale_factor=self.opt.hw_scale, mode='bilinear')         # :287
        sil = F.interpolate(sil, scale_factor=self.opt.hw_scale, mode='bilinear')             # :288
        if itr > self.start_fusion_step:                                                      # :294
            with torch.no_grad():
                latents = self.vae.encode(normalize(image)).mode() * self.z_scale_factor      # :299
                if max_thres is None:
                    max_thres = u.clamp(min=0.0, max=0.99).item()                             # :303
                pred_x0, x_noisy, noise, alpha_cumprod = self.sampler.sample(latents, cond_images=feats, use_tqdm=False, return_noise=True,
                                                                             max_thres=max_thres)            # :304
                fusion_weight = (1 - alpha_cumprod).to(pred_x0.device)                        # :307
                pred_img = unnormalize(self.vae.decode(1.0 / self.z_scale_factor * pred_x0)).clip(0.0, 1.0)   # :309
            fusion_loss = (fusion_weight * (image - pred_img).abs().mean()).sum()             # :310 ([1]-shaped in the reference)
            if self.percep is not None and itr >= self.start_percep_step:                     # :176-178, :312-314
                fusion_loss = fusion_loss + self.percep(image, pred_img, normalize=True).mean() * self.lambda_percep
            self.last['unet_calls'] = self.sampler.last_unet_calls
        else:                                                                                 # EFT bootstrap :316-329
            noisy_rgb = c.target_eft_image[vi:vi + 1]
            noisy_mask = (noisy_rgb.mean(dim=1, keepdim=True) > .1).float()                   # :269-271
            fusion_loss = self.lambda_color * huber(image, noisy_rgb).abs().mean() + self.lambda_sil * huber(sil, noisy_mask).abs().mean()
        opacity = torch.sqrt(sil ** 2 + .01).mean()                                           # :336
        loss = fusion_loss + self.lambda_opacity * opacity                                    # :344
        loss.backward()                                                                       # :345
        self.optimizer.step(grad_scale=self.optimizer.sync_grads(self.world_size, self.pg))   # :352
        self.last['fusion_loss'] = loss.detach()
        return loss.detach()

    def _fusion_substep_fused(self, itr, max_thres, vi, u, feats):
        """the same sub-step with the image-space work in three kernels: bilinear x2 to NCHW (for the VAE), loss value + gradient at full
        resolution, adjoint of the bilinear back to the render's [N,3] / [N] outputs"""
        c = self.cache
        img, ws, hw = self._render_raw(c.target_rays_o[vi], c.target_rays_d[vi], 'B')
        up = glue.upsample2x_render(img.detach(), ws.detach(), hw, hw)                         # :287-288, planes (r, g, b, opacity)
        if itr > self.start_fusion_step:                                                      # :294
            with torch.no_grad():
                latents = self.vae.encode(normalize(up[None, :3])).mode() * self.z_scale_factor          # :299
                if max_thres is None:
                    max_thres = u.clamp(min=0.0, max=0.99).item()                             # :303
                pred_x0, _, _, _ = self.sampler.sample(latents, cond_images=feats, use_tqdm=False, return_noise=True, max_thres=max_thres)   # :304
                pred_img = unnormalize(self.vae.decode(1.0 / self.z_scale_factor * pred_x0)).clip(0.0, 1.0)   # :309
                if self.pred_img_hook is not None:
                    pred_img = self.pred_img_hook([vi], pred_img)
            weight = 1.0 - plms_sigmoid(plms_log_snr(float(max_thres)))                      # 1 - alpha_cumprod of the run's first noise level (:307)
            g_extra, percep_term = None, None
            if self.percep is not None and itr >= self.start_percep_step:                     # :176-178, :312-314: value and gradient together
                pv, pg = self.percep.value_and_grad(up[:3], pred_img[0], normalize=True)
                percep_term, g_extra = pv * self.lambda_percep, pg * self.lambda_percep
            loss, g_img, g_ws = glue.fusion_loss(up, pred_img[0], hw, hw, 'sds', weight, self.lambda_color, self.lambda_sil, self.lambda_opacity,
                                                 g_extra=g_extra)
            if percep_term is not None:
                loss = loss + percep_term
            self.last['unet_calls'] = self.sampler.last_unet_calls
        else:                                                                                 # EFT bootstrap :316-329
            loss, g_img, g_ws = glue.fusion_loss(up, c.target_eft_image[vi], hw, hw, 'eft', 1.0, self.lambda_color, self.lambda_sil, self.lambda_opacity)
        torch.autograd.backward((img, ws), (g_img, g_ws))                                     # :345
        self.optimizer.step(grad_scale=self.optimizer.sync_grads(self.world_size, self.pg))   # :352
        self.last['fusion_loss'] = loss
        return loss

    def step(self, itr: int, max_thres: Optional[float] = None):
        """one iteration of the reference's main loop (distillation.py:174-352); returns the two losses as device scalars"""
        with torch.cuda.device(self.cache.input_rgb.device):
            if self.views_per_step is not None:
                return self.minibatch_step(itr, max_thres)
            a = self.photometric_substep(itr)
            b = self.fusion_substep(itr, max_thres)
        return a, b

    # ------------------------------------------------------------------------------------------------ view-batched step
    def _keyed_render(self, rays_o, rays_d, itr, view, stream):
        """render with this (iteration, view)'s own jitter / inverse-CDF draws"""
        n = rays_o.shape[0]
        if self.render_noise_keyed is not None:
            pn, un = self.render_noise_keyed(('A', 'B')[stream], view)
        else:
            g = torch.Generator(device=rays_o.device)
            g.manual_seed(view_seed(self.seed, itr, view, stream))
            pn = torch.rand(n, self.opt.num_steps, generator=g, device=rays_o.device)            # renderer_df.py:363
            un = torch.rand(n, self.opt.upsample_steps, generator=g, device=rays_o.device)       # renderer_df.py:31
        out = self.ngp.render(rays_o[None], rays_d[None], **dict(self.render_kw, perturb_noise=pn, pdf_noise=un))
        return out['image'].reshape(n, 3), out['weights_sum'].reshape(n), int(round(n ** 0.5))

    def minibatch_step(self, itr: int, max_thres: Optional[float] = None):
        """one optimiser step on (photometric view + V target views); returns (photometric loss, mean fusion loss) as device scalars.  With
        world_size R each rank handles views[r::R]; the losses returned are the GLOBAL ones only on the rank(s) that computed every term --
        use them for logging, not for control flow."""
        if not self.fused_glue or int(self.opt.hw_scale) != 2:
            raise NotImplementedError('the view-batched step uses the fused image-space kernels (fused_glue=True, hw_scale=2)')
        c = self.cache
        V, R, r = int(self.views_per_step), self.world_size, self.rank
        dev = c.input_rgb.device
        n_in, n_t = c.input_rgb.shape[0], c.target_features.shape[0]
        idx = int(torch.randperm(n_in, generator=self.gen)[0])                                 # :185-186
        perm = torch.randperm(n_t, generator=self.gen)                                         # :263
        u = torch.rand(1, generator=self.gen)                                                  # :303: ONE noise level per step, shared by its views
        views = step_views(perm, V)
        mine = shard_views(views, r, R)
        self.ngp.train()
        self.optimizer.zero_grad()
        zero = torch.zeros((), device=dev)
        loss_a = zero
        # ---- photometric term: one rank (round-robin over iterations), gradient joins the step's single all-reduce
        if itr % R == r:
            if self.opt.cuda_ray and itr % 16 == 0:
                self.ngp.update_extra_state()
            img, ws, hw = self._keyed_render(c.input_rays_o[idx], c.input_rays_d[idx], itr, idx, 0)
            loss_a, g_img, g_ws = glue.photometric_loss(img.detach(), ws.detach(), c.input_rgb[idx], c.input_mask[idx], hw, hw, int(self.opt.hw_scale),
                                                        self.lambda_color, self.lambda_sil, self.lambda_opacity)
            torch.autograd.backward((img, ws), (g_img, g_ws))
        elif self.opt.cuda_ray and itr % 16 == 0:
            self.ngp.update_extra_state()
        # ---- fusion term over this rank's views, in batches of at most max_batch
        sds = itr > self.start_fusion_step
        if sds and max_thres is None:
            max_thres = u.clamp(min=0.0, max=0.99).item()
        loss_b = zero
        calls = 0
        inv_v = 1.0 / V
        for lo in range(0, len(mine), self.max_batch):
            chunk = mine[lo:lo + self.max_batch]
            renders = [self._keyed_render(c.target_rays_o[v], c.target_rays_d[v], itr, v, 1) for v in chunk]
            hw = renders[0][2]
            ups = [glue.upsample2x_render(im.detach(), w.detach(), hw, hw) for im, w, _ in renders]          # :287-288
            if sds:
                with torch.no_grad():
                    batch = torch.stack([up[:3] for up in ups])                                               # [b,3,256,256]
                    latents = self.vae.encode(normalize(batch)).mode() * self.z_scale_factor                 # :299
                    feats = c.target_features[chunk] if isinstance(c.target_features, torch.Tensor) else torch.cat([c.target_features[v:v + 1] for v in chunk])
                    prev_fn = self.sampler.noise_fn
                    if self.plms_noise is None:
                        self.sampler.noise_fn = KeyedNoise([view_seed(self.seed, itr, v, 2) for v in chunk], dev)
                    else:
                        self.sampler.noise_fn = self.plms_noise(chunk)
                    try:
                        pred_x0, _, _, _ = self.sampler.sample(latents, cond_images=feats, use_tqdm=False, return_noise=True, max_thres=max_thres)   # :304
                    finally:
                        self.sampler.noise_fn = prev_fn
                    pred_img = unnormalize(self.vae.decode(1.0 / self.z_scale_factor * pred_x0)).clip(0.0, 1.0)   # :309
                    if self.pred_img_hook is not None:
                        pred_img = self.pred_img_hook(chunk, pred_img)
                calls = self.sampler.last_unet_calls
                weight = 1.0 - plms_sigmoid(plms_log_snr(float(max_thres)))                                  # :307
            for j, v in enumerate(chunk):
                im, w, _ = renders[j]
                if sds:
                    g_extra, percep_term = None, None
                    if self.percep is not None and itr >= self.start_percep_step:                             # :176-178, :312-314
                        pv, pg = self.percep.value_and_grad(ups[j][:3], pred_img[j], normalize=True)
                        percep_term, g_extra = pv * self.lambda_percep, pg * self.lambda_percep
                    lv, g_img, g_ws = glue.fusion_loss(ups[j], pred_img[j], hw, hw, 'sds', weight, self.lambda_color, self.lambda_sil,
                                                       self.lambda_opacity, g_extra=g_extra)
                    if percep_term is not None:
                        lv = lv + percep_term
                else:
                    lv, g_img, g_ws = glue.fusion_loss(ups[j], c.target_eft_image[v], hw, hw, 'eft', 1.0, self.lambda_color, self.lambda_sil,
                                                       self.lambda_opacity)
                torch.autograd.backward((im, w), (g_img * inv_v, g_ws * inv_v))                               # mean over the step's V views
                loss_b = loss_b + lv * inv_v
            del renders, ups
        if sds:
            self.last['unet_calls'] = calls
        self.optimizer.sync_grads(R, self.pg)              # the step's ONE collective: sum of every rank's (photometric + fusion) gradient
        self.optimizer.step(grad_scale=1.0)
        self.optimizer.scheduler_step()                                                        # :247
        self.last['photo_loss'], self.last['fusion_loss'] = loss_a, loss_b
        self.last['views'] = mine
        return loss_a, loss_b


# ----------------------------------------------------------------------------------------------------------------------
# distillation_loop: the reference's entry point (sparsefusion/distillation.py:26-497), same signature
# ----------------------------------------------------------------------------------------------------------------------
from .network_grid import NeRFNetwork, get_default_torch_ngp_opt  # noqa: E402,F401  (re-exported: demo.py:9 imports both names from this module)


class ReferenceSceneBuilder:
    """distillation.py:60-127 -- what the reference prepares per scene before its loop: relative / voxel-frame cameras, 50 circle-path
    augmentation cameras, the 128x128 ray sampler, and the EFT feature cache of every (scene + augmentation) view.  These steps are outside the
    hot path: they are executed by the REFERENCE's own utilities -- ``utils.camera_utils``, ``utils.render_utils`` (pytorch3d)
    and the EFT model handed in through ``model_tuple`` -- imported at call time from the reference tree the caller runs in.  The result is the
    ``SceneCache`` the hot loop consumes: rays as tensors, cached features / low-res EFT images per target view."""

    def __init__(self, gpu):
        self.gpu = gpu

    def build(self, eft, scene_cameras, scene_rgb, scene_mask, input_idx, use_diffusion=True) -> 'SceneCache':
        try:
            from utils.camera_utils import RelativeCameraLoader, get_interpolated_path
            from utils.render_utils import init_light_field_renderer, init_ray_sampler
        except ImportError as e:
            raise RuntimeError('distillation_loop needs the reference tree (utils/camera_utils.py, utils/render_utils.py) and pytorch3d on sys.path to turn '
                               'pytorch3d cameras into rays and to run the EFT; pass scene_builder=... to supply a SceneCache another way') from e
        from einops import rearrange
        gpu = self.gpu
        relative_cam = RelativeCameraLoader(relative=True, center_at_origin=True)                                    # :65-66
        relative_cam_no_origin = RelativeCameraLoader(relative=True, center_at_origin=False)
        scene_cameras_vox = relative_cam_no_origin.get_relative_camera(scene_cameras, query_idx=[0], center_at_origin=False)   # :70
        scene_cameras_aug = get_interpolated_path(scene_cameras, n=50, method='circle', theta_offset_max=0.17)      # :73
        scene_cameras_aug = relative_cam.concat_cameras([scene_cameras, scene_cameras_aug])                          # :74
        scene_cameras_aug_rel = relative_cam.get_relative_camera(scene_cameras_aug, query_idx=[0], center_at_origin=True)      # :75
        scene_cameras_aug_vox = relative_cam_no_origin.get_relative_camera(scene_cameras_aug, query_idx=[0], center_at_origin=False)
        blank_rgb = torch.zeros_like(scene_rgb[:1]).repeat(len(scene_cameras_aug), 1, 1, 1)                          # :77-79
        scene_rgb_aug = torch.cat((scene_rgb, blank_rgb))
        cam_dist_mean = torch.mean(torch.linalg.norm(scene_cameras.get_camera_center(), axis=1))                     # :82-84
        min_depth, volume_extent_world = cam_dist_mean - 5.0, cam_dist_mean + 5.0
        _, _, sampler_feat = init_ray_sampler(gpu, 256, 256, min=min_depth, max=volume_extent_world, scale_factor=2)            # :85
        _, _, renderer_feat = init_light_field_renderer(gpu, 256, 256, min=min_depth, max=volume_extent_world, scale_factor=8.0)   # :86

        def rays_of(cam):
            rb = sampler_feat(cam)                                                                                  # :201-204, :274-277
            return rearrange(rb.origins, 'b h w c -> b (h w) c')[0].contiguous(), rearrange(rb.directions, 'b h w c -> b (h w) c')[0].contiguous()
        in_o, in_d = zip(*[rays_of(relative_cam.get_camera_slice(scene_cameras_vox, [i])) for i in input_idx])
        feats, eft_imgs, t_o, t_d = [], [], [], []
        n_cache = len(scene_cameras_aug_rel) if use_diffusion else 0
        for ci in range(n_cache):                                                                                   # :95-125
            q_cam, _, _, input_cameras, input_rgb, _ = relative_cam(scene_cameras_aug_rel, scene_rgb_aug, query_idx=[ci], context_idx=input_idx)
            eft.encode(input_cameras, input_rgb)
            with torch.no_grad():
                epi, _, _ = renderer_feat(cameras=q_cam, volumetric_function=eft.batched_forward, n_batches=16, input_cameras=input_cameras, input_rgb=input_rgb)
                lr_render, latents = epi.split([3, 256], dim=-1)
            feats.append(rearrange(latents, 'b h w f -> b f h w'))
            eft_imgs.append(F.interpolate(rearrange(lr_render, 'b h w f -> b f h w'), scale_factor=8.0, mode='bilinear'))
            o, d = rays_of(relative_cam.get_camera_slice(scene_cameras_aug_vox, [ci]))
            t_o.append(o)
            t_d.append(d)
        dev = scene_rgb.device
        n_rays = in_o[0].shape[0]
        empty = lambda *s: torch.zeros(0, *s, device=dev)
        return SceneCache(input_rgb=scene_rgb[input_idx].float(), input_mask=scene_mask[input_idx].float(), input_rays_o=torch.stack(in_o), input_rays_d=torch.stack(in_d),
                          target_features=torch.cat(feats) if feats else empty(256, 32, 32), target_eft_image=torch.cat(eft_imgs) if eft_imgs else empty(3, 256, 256),
                          target_rays_o=torch.stack(t_o) if t_o else empty(n_rays, 3), target_rays_d=torch.stack(t_d) if t_d else empty(n_rays, 3))


def distillation_loop(gpu, args, opt, model_tuple, save_dir, seq_name, scene_cameras, scene_rgb, scene_mask, scene_valid_region, input_idx,
                      use_diffusion=True, max_itr=3000, loss_fn_vgg=None, *, scene_builder=None, views_per_step=None, on_iteration=None, seed=None,
                      process_group=None):
    """sparsefusion/distillation.py:26-497 with the reference's signature (demo.py:87-103 calls it unchanged): builds the per-scene cache
    (``ReferenceSceneBuilder``, or ``scene_builder.build(...)``), optimises a fresh ``NeRFNetwork(opt)`` for ``max_itr`` iterations with
    ``Distiller.step`` and saves ``{'model_state_dict': ngp.state_dict()}`` to ``{save_dir}/{seq_name}.pt`` (:495-496).

    Not reproduced: the matplotlib loss plots and the periodic / final visualisation renders, gifs and debug metrics (:355-388, :391-493) -- logging,
    outside the hot path; ``on_iteration(itr, distiller)`` is the hook for them (``distiller.ngp.render_batched`` is the same call the reference
    makes).  ``scene_valid_region`` and ``loss_fn_vgg`` are accepted and unused, as in the reference's loop body (:192-195 overwrites the region with
    ones; loss_fn_vgg only feeds the debug metrics).  Extensions, keyword-only: ``views_per_step`` (view-batched step, see the module docstring; with
    torch.distributed initialised and ``process_group`` given the views are sharded over its ranks), ``seed``.
    Returns the optimised network."""
    import os
    eft, vae, vldm = model_tuple
    exp_dir = getattr(args, 'exp_dir', save_dir)
    os.makedirs(f'{exp_dir}/render_imgs/{seq_name}/', exist_ok=True)                                               # :57-58
    os.makedirs(f'{exp_dir}/render_gifs/', exist_ok=True)
    os.makedirs(save_dir, exist_ok=True)
    device = torch.device('cuda', gpu) if isinstance(gpu, int) else torch.device(gpu)
    builder = scene_builder if scene_builder is not None else ReferenceSceneBuilder(gpu)
    cache = builder.build(eft, scene_cameras, scene_rgb, scene_mask, input_idx, use_diffusion=use_diffusion)
    from .lpips_vgg import PerceptualLoss
    with torch.cuda.device(device):
        perceptual_loss = PerceptualLoss('vgg', device=device) if use_diffusion else None                          # :161
        ngp_network = NeRFNetwork(opt).to(device).train()                                                          # :164
        rank, world = 0, 1
        if process_group is not None:
            rank, world = torch.distributed.get_rank(process_group), torch.distributed.get_world_size(process_group)
        dist = Distiller(ngp_network, vae, vldm, opt, cache.to(device), plms_steps=50, start_fusion_step=1000, lambda_color=1.0, lambda_sil=1.0,   # :148-160
                         lambda_opacity=1e-3, seed=torch.seed() % (2 ** 31) if seed is None else seed, rank=rank, world_size=world,
                         process_group=process_group, percep=perceptual_loss, lambda_percep=0.1, start_percep_step=1000, views_per_step=views_per_step)
        for itr in range(max_itr):                                                                                  # :174
            if use_diffusion:
                dist.step(itr)
            else:
                dist.photometric_substep(itr)
            if on_iteration is not None:
                on_iteration(itr, dist)
        torch.cuda.synchronize(device)
    w_addr = f'{save_dir}/{seq_name}.pt'                                                                            # :495-496
    torch.save({'model_state_dict': ngp_network.state_dict()}, w_addr)
    print('input idx', input_idx)
    return ngp_network
