"""PLMSSampler -- host-side mirror of external/plms.py (reference): the pseudo-linear-multistep sampler the
distillation loop calls once per SDS step (sparsefusion/distillation.py:160,304).

Same constructor and ``sample(image, max_thres, cond_images, cond_scale, use_tqdm, return_noise)`` signature and
return values ``(pred_x0, x_noisy, noise, alpha_cumprod)``.  ``n_steps = min(int(max_thres*plms_steps*2), plms_steps)``
steps, ``n_steps + 1`` UNet evaluations (plms.py:87,136-142), Adams-Bashforth 2/3/4 on the eps history (:144-152),
clamp of x0 to +-clip_value (:205), stochastic posterior step (:208-212).

B200 notes: each eps evaluation replays one CUDA graph of the UNet (``UnetGraph``); the few-hundred-byte schedule
math between evaluations stays in torch.  Every ``torch.randn_like`` of the reference is routed through
``self.noise_fn`` (default ``torch.randn_like``) in the reference's draw order, so parity tests can inject the
oracle's noise.
"""
from __future__ import annotations

import torch

from .imagen_pytorch import GaussianDiffusionContinuousTimes, UnetGraph


class PLMSSampler:
    def __init__(self, diffusion, plms_steps=100, use_cuda_graph: bool = True, noise_fn=None):
        self.diffusion = diffusion
        self.plms_steps = plms_steps
        self.noise_fn = noise_fn if noise_fn is not None else torch.randn_like
        self.use_cuda_graph = use_cuda_graph
        self._graph = None
        self.last_unet_calls = 0

    def _eps(self, unet, x, log_snr, cond_images, cond_scale):
        self.last_unet_calls += 1
        if self.use_cuda_graph and cond_scale == 1:
            if self._graph is None or self._graph.unet is not unet:
                self._graph = UnetGraph(unet)
            return self._graph(x, log_snr, cond_images).clone()
        return unet.forward_with_cond_scale(x, log_snr, cond_images=cond_images, cond_scale=cond_scale)

    @torch.no_grad()
    def sample(self, image=None, max_thres=.999, cond_images=None, cond_scale=1.0, use_tqdm=True, return_noise=False, **kwargs):
        batch_size = cond_images.shape[0]
        d = self.diffusion
        shape = (batch_size, d.sample_channels[0], d.image_sizes[0], d.image_sizes[0])
        img, x_noisy, noise, alpha_cumprod = self.plms_sample_loop(d.unets[0], image=image, shape=shape, cond_images=cond_images,
                                                                   cond_scale=cond_scale, noise_scheduler=d.noise_schedulers[0],
                                                                   pred_objective=d.pred_objectives[0],
                                                                   dynamic_threshold=d.dynamic_thresholding[0], use_tqdm=use_tqdm,
                                                                   max_thres=max_thres)
        if not return_noise:
            return img
        return img, x_noisy, noise, alpha_cumprod

    @torch.no_grad()
    def plms_sample_loop(self, unet, image, shape, cond_images, cond_scale, noise_scheduler, pred_objective, dynamic_threshold,
                         use_tqdm, max_thres=None):
        batch, device = shape[0], self.diffusion.device
        self.last_unet_calls = 0
        if image is None:
            image = torch.randn(shape, device=device)
        else:
            assert max_thres is not None
        short = GaussianDiffusionContinuousTimes(noise_schedule='cosine', timesteps=self.plms_steps)
        if max_thres >= .99:
            timesteps = short.get_sampling_timesteps(batch, device=device)
            noise = self.noise_fn(image)
            x_noisy, log_snr = short.q_sample(image, t=max_thres, noise=noise)
            img = image
        else:
            n_steps = min(int(max_thres * self.plms_steps * 2), self.plms_steps)
            timesteps = short.get_sampling_timesteps_custom(batch, device=device, max_thres=max_thres, n_steps=n_steps)
            noise = self.noise_fn(image)
            img, log_snr = short.q_sample(image, t=max_thres, noise=noise)
            x_noisy = img
        old_eps = []
        for times, times_next in timesteps:
            img, pred_x0, e_t = self.p_sample(unet, img, times, t_next=times_next, cond_images=cond_images, cond_scale=cond_scale,
                                              noise_scheduler=noise_scheduler, pred_objective=pred_objective,
                                              dynamic_threshold=dynamic_threshold, old_eps=old_eps)
            old_eps.append(e_t)
            if len(old_eps) >= 4:
                old_eps.pop(0)
        if self.diffusion.clip_output:
            img = img.clamp(-self.diffusion.clip_value, self.diffusion.clip_value)
        return self.diffusion.unnormalize_img(img), x_noisy, noise, torch.sigmoid(log_snr)

    @torch.no_grad()
    def p_sample(self, unet, x, t, t_next, cond_images, cond_scale, noise_scheduler, pred_objective, dynamic_threshold, old_eps):
        args = (cond_images, cond_scale, noise_scheduler, pred_objective, dynamic_threshold)
        _, _, e_t = self.get_model_output(unet, x, t, t_next, *args)
        if len(old_eps) == 0:
            x_prev, _, _ = self.get_model_output(unet, x, t, t_next, *args, pred_e=e_t)
            _, _, e_t_next = self.get_model_output(unet, x_prev, t_next, t_next, *args)
            e_t_prime = (e_t + e_t_next) / 2
        elif len(old_eps) == 1:
            e_t_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_t_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_t_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        x_prev, pred_x0, _ = self.get_model_output(unet, x, t, t_next, *args, pred_e=e_t_prime)
        return x_prev, pred_x0, e_t

    def get_model_output(self, unet, x, t, t_next, cond_images, cond_scale, noise_scheduler, pred_objective, dynamic_threshold, pred_e=None):
        assert pred_objective == 'noise'
        b = x.shape[0]
        if pred_e is None:
            pred_e = self._eps(unet, x, noise_scheduler.get_condition(t), cond_images, cond_scale)
        x_start = noise_scheduler.predict_start_from_noise(x, t=t, noise=pred_e)
        if self.diffusion.clip_output:
            if dynamic_threshold:
                raise NotImplementedError('dynamic thresholding is unreachable in the reference too (plms.py:198 reads a missing attribute)')
            x_start = x_start.clamp(-self.diffusion.clip_value, self.diffusion.clip_value)
        model_mean, _, model_log_variance = noise_scheduler.q_posterior(x_start=x_start, x_t=x, t=t, t_next=t_next)
        noise = self.noise_fn(x)
        nonzero_mask = (1 - (t_next == 0).float()).reshape(b, *((1,) * (len(x.shape) - 1)))
        x_prev = model_mean + nonzero_mask * (0.5 * model_log_variance).exp() * noise
        return x_prev, x_start, pred_e
