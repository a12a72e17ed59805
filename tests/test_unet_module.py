"""Host-side Unet / DDPM mirror: state_dict inventory equals the reference's (through the oracle's inventory, which
oracle/gen_golden.py checked key-for-key against external.imagen_pytorch.Unet), and it refuses to run without CUDA."""
import numpy as np
import pytest
import torch


def _kwargs(cfg):
    return dict(channels=cfg.channels, dim=cfg.dim, dim_mults=cfg.dim_mults, num_resnet_blocks=cfg.num_resnet_blocks,
                layer_attns=cfg.layer_attns, layer_cross_attns=tuple(False for _ in cfg.dim_mults),
                cond_images_channels=cfg.cond_images_channels, attn_pool_text=False, attn_dim_head=cfg.attn_dim_head,
                attn_heads=cfg.attn_heads)


def test_small_unet_state_dict_matches_reference_inventory():
    from oracle import unet_oracle as uo
    from sparsefusion_b200.imagen_pytorch import Unet
    from sparsefusion_b200.vldm import DDPM
    unet = Unet(**_kwargs(uo.SMALL))            # as utils/load_model.py:60-69 builds it (cond_on_z defaults to True)
    assert 'conditional_to_cond.weight' in unet.state_dict()
    ddpm = DDPM(channels=4, unets=(unet,), conditional_encoder=None, conditional_embed_dim=None, image_sizes=(uo.SMALL.image_size,),
                timesteps=500, cond_drop_prob=0.1, pred_objectives='noise', conditional=False, auto_normalize_img=False,
                clip_output=True, dynamic_thresholding=False, dynamic_thresholding_percentile=.68, clip_value=10)
    live = ddpm.unets[0]
    assert live is not unet and live.cond_on_z is False   # re-instantiated like sparsefusion/vldm.py:165-171
    want = uo.param_shapes(uo.SMALL)
    got = {k: tuple(v.shape) for k, v in live.state_dict().items()}
    assert got == want
    assert all(k.startswith('unets.0.') for k in ddpm.state_dict().keys())
    sd = uo.make_params(uo.SMALL)
    live.load_state_dict(sd, strict=True)
    assert torch.equal(live.state_dict()['final_conv.weight'], sd['final_conv.weight'])
    assert (Unet(**_kwargs(uo.SMALL)).state_dict()['final_conv.weight'] == 0).all()   # zero-init like imagen_pytorch.py:1388
    with pytest.raises(RuntimeError, match='CUDA'):
        live.forward(torch.zeros(1, 4, 16, 16), torch.zeros(1), cond_images=torch.zeros(1, 12, 16, 16))


def test_full_inventory_is_477_tensors():
    from oracle import unet_oracle as uo
    from sparsefusion_b200.imagen_pytorch import unet_param_shapes
    cfg = uo.FULL
    s = unet_param_shapes(dim=cfg.dim, dim_mults=cfg.dim_mults, num_resnet_blocks=cfg.num_resnet_blocks, layer_attns=cfg.layer_attns,
                          channels=4, channels_out=4, cond_images_channels=256, cond_dim=256, attn_dim_head=64, attn_heads=8, ff_mult=2.,
                          num_time_tokens=2, learned_sinu_pos_emb_dim=16, init_cross_embed_kernel_sizes=(3, 7, 15), max_conditional_len=256,
                          cond_on_z=False, conditional_embed_dim=None)
    assert s == uo.param_shapes(cfg) and len(s) == 477


def test_unsupported_configurations_fail_loudly():
    from sparsefusion_b200.imagen_pytorch import Unet
    with pytest.raises(NotImplementedError):
        Unet(dim=32, channels=4, layer_cross_attns=True, attn_pool_text=False)
    with pytest.raises(NotImplementedError):
        Unet(dim=32, channels=4, layer_cross_attns=False, attn_pool_text=False, memory_efficient=True)


def test_schedule_math_matches_oracle():
    from oracle import unet_oracle as uo
    from sparsefusion_b200.imagen_pytorch import GaussianDiffusionContinuousTimes
    ns = GaussianDiffusionContinuousTimes(noise_schedule='cosine', timesteps=50)
    t = torch.tensor([0.0, 0.02, 0.37, 0.98])
    assert torch.equal(ns.get_condition(t), uo.alpha_cosine_log_snr(t))
    x, n = torch.randn(4, 4, 8, 8), torch.randn(4, 4, 8, 8)
    assert torch.equal(ns.q_sample(x, t, n)[0], uo.q_sample(x, t, n)[0])
    assert torch.equal(ns.predict_start_from_noise(x, t, n), uo.predict_start_from_noise(x, t, n))
    tn = (t - 0.01).clamp(min=0)
    for a, b in zip(ns.q_posterior(x, n, t, t_next=tn), uo.q_posterior(x, n, t, tn)):
        assert torch.equal(a, b)
    pairs = ns.get_sampling_timesteps_custom(2, device='cpu', max_thres=0.37, n_steps=37)
    assert len(pairs) == 37 and pairs[-1][1].eq(0).all() and pairs[0][0].shape == (2,)
