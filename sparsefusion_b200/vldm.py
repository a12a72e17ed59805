"""DDPM -- host-side mirror of sparsefusion/vldm.py:53-285 (reference): the container the SparseFusion pipeline
builds around the VLDM UNet (utils/load_model.py:76-91) and hands to PLMSSampler (sparsefusion/distillation.py:160).

Kept: constructor keywords, ``unets`` (re-instantiated through ``Unet.cast_model_parameters`` exactly like
vldm.py:165-171, so checkpoints with keys ``unets.0.*`` load), ``noise_schedulers``, ``pred_objectives``,
``dynamic_thresholding``, ``sample_channels``, ``image_sizes``, ``clip_output``, ``clip_value``,
``unnormalize_img`` / ``normalize_img``, ``device``, ``state_dict`` / ``load_state_dict``.
Not kept: the training loss (``forward``, vldm.py:711-776) and ancestral ``sample`` (:445-555) -- training the
diffusion model is outside the distillation hot path; they raise NotImplementedError.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .imagen_pytorch import GaussianDiffusionContinuousTimes, Unet, cast_tuple


def _identity(t, *a, **k):
    return t


class DDPM(nn.Module):
    def __init__(self, unets, *, image_sizes, conditional_encoder, conditional_embed_dim=1024, channels=3, timesteps=1000,
                 cond_drop_prob=0.1, loss_type='l2', noise_schedules='cosine', pred_objectives='noise', random_crop_sizes=None,
                 lowres_noise_schedule='linear', lowres_sample_noise_level=0.2, per_sample_random_aug_noise_level=False, conditional=True,
                 auto_normalize_img=False, p2_loss_weight_gamma=0.5, p2_loss_weight_k=1, dynamic_thresholding=True,
                 dynamic_thresholding_percentile=0.95, only_train_unet_number=None, clip_output=True, clip_value=1.0):
        super().__init__()
        self.timesteps = timesteps
        self.loss_type = loss_type
        self.conditional, self.unconditional = conditional, not conditional
        self.channels = channels
        unets = cast_tuple(unets)
        num_unets = len(unets)
        if num_unets != 1:
            raise NotImplementedError('cascaded DDPMs are not part of the SparseFusion path (one base unet: utils/load_model.py:76)')
        timesteps = cast_tuple(timesteps, num_unets)
        noise_schedules = cast_tuple(noise_schedules)
        noise_schedules = (*noise_schedules, *(('cosine',) * max(0, 2 - len(noise_schedules))))[:max(2, num_unets)]
        self.noise_schedulers = nn.ModuleList([GaussianDiffusionContinuousTimes(noise_schedule=ns, timesteps=ts)
                                               for ts, ns in zip(timesteps, noise_schedules)])
        self.lowres_noise_schedule = GaussianDiffusionContinuousTimes(noise_schedule=lowres_noise_schedule)
        self.pred_objectives = cast_tuple(pred_objectives, num_unets)
        self.conditional_embed_dim = conditional_embed_dim
        self.encode_conditional = conditional_encoder
        self.unets = nn.ModuleList([])
        self.unet_being_trained_index = -1
        self.only_train_unet_number = only_train_unet_number
        for ind, one_unet in enumerate(unets):
            assert isinstance(one_unet, Unet)
            one_unet = one_unet.cast_model_parameters(lowres_cond=not (ind == 0), cond_on_z=self.conditional,
                                                      conditional_embed_dim=self.conditional_embed_dim if self.conditional else None,
                                                      channels=self.channels, channels_out=self.channels)
            self.unets.append(one_unet)
        self.image_sizes = cast_tuple(image_sizes)
        assert num_unets == len(self.image_sizes)
        self.sample_channels = cast_tuple(self.channels, num_unets)
        self.is_video = False
        self.lowres_sample_noise_level = lowres_sample_noise_level
        self.per_sample_random_aug_noise_level = per_sample_random_aug_noise_level
        self.cond_drop_prob = cond_drop_prob
        self.can_classifier_guidance = cond_drop_prob > 0.
        self.normalize_img = (lambda img: img * 2 - 1) if auto_normalize_img else _identity
        self.unnormalize_img = (lambda img: (img + 1) * 0.5) if auto_normalize_img else _identity
        self.input_image_range = (0. if auto_normalize_img else -1., 1.)
        self.dynamic_thresholding = cast_tuple(dynamic_thresholding, num_unets)
        self.dynamic_thresholding_percentile = dynamic_thresholding_percentile
        self.clip_output, self.clip_value = clip_output, clip_value
        self.p2_loss_weight_k = p2_loss_weight_k
        self.p2_loss_weight_gamma = cast_tuple(p2_loss_weight_gamma, num_unets)
        self.register_buffer('_temp', torch.tensor([0.]), persistent=False)
        self.to(next(self.unets.parameters()).device)

    @property
    def device(self):
        return self._temp.device

    def get_unet(self, unet_number):
        assert 0 < unet_number <= len(self.unets)
        return self.unets[unet_number - 1]

    def reset_unets_all_one_device(self, device=None):
        self.unets.to(device if device is not None else self.device)

    def forward(self, *a, **k):
        raise NotImplementedError('DDPM.forward is the VLDM training loss (sparsefusion/vldm.py:711-776): outside the distillation hot path')

    def sample(self, *a, **k):
        raise NotImplementedError('ancestral sampling (sparsefusion/vldm.py:445-555) is outside the distillation hot path; use PLMSSampler')
