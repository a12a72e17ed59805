"""The UNet / PLMS restatement (oracle/unet_oracle.py) reproduces the REFERENCE's own outputs stored in
tests/golden/unet_*.npz and plms_small.npz (minted by oracle/gen_golden.py from external.imagen_pytorch.Unet,
sparsefusion.vldm.DDPM and external.plms.PLMSSampler).  Tolerance: fp32 CPU, same op sequence -> 1e-5 rel."""
import numpy as np
import pytest
import torch


def _inputs(cfg, batch, seed):
    rng = np.random.default_rng(seed)
    h = cfg.image_size
    x = torch.from_numpy(rng.standard_normal((batch, cfg.channels, h, h), dtype=np.float32))
    cond = torch.from_numpy(rng.standard_normal((batch, cfg.cond_images_channels, h, h), dtype=np.float32))
    return x, cond


def test_param_inventory_matches_reference_state_dict_size():
    from oracle import unet_oracle as uo
    shapes = uo.param_shapes(uo.FULL)
    assert len(shapes) == 477                                   # 477 tensors under unets.0.
    assert sum(int(np.prod(s)) for s in shapes.values()) == 400_675_357


def test_small_unet_matches_reference_golden(golden_dir):
    from oracle import unet_oracle as uo
    g = np.load(f'{golden_dir}/unet_small.npz')
    sd = uo.make_params(uo.SMALL, seed=int(g['seed_params']))
    x, cond = _inputs(uo.SMALL, int(g['batch']), int(g['seed_inputs']))
    taps = {}
    with torch.no_grad():
        eps = uo.unet_forward(sd, uo.SMALL, x, uo.alpha_cosine_log_snr(torch.from_numpy(g['t'])), cond, taps)
    ref = torch.from_numpy(g['eps'])
    assert ((eps - ref).norm() / ref.norm()).item() < 1e-5
    for k in g.files:
        if k.startswith('tap_'):
            t = torch.from_numpy(g[k])
            assert ((taps[k[4:]] - t).norm() / t.norm()).item() < 1e-5, k


def test_unet_at_128x128_latents_matches_reference_golden(golden_dir):
    """BASELINE configs[4] geometry: 128x128x4 latents (256 queries x 259 keys at the attention stage), narrow width"""
    import dataclasses
    from oracle import unet_oracle as uo
    cfg = dataclasses.replace(uo.SMALL, image_size=128)
    g = np.load(f'{golden_dir}/unet_small128.npz')
    sd = uo.make_params(cfg, seed=int(g['seed_params']))
    x, cond = _inputs(cfg, int(g['batch']), int(g['seed_inputs']))
    taps = {}
    with torch.no_grad():
        eps = uo.unet_forward(sd, cfg, x, uo.alpha_cosine_log_snr(torch.from_numpy(g['t'])), cond, taps)
    ref = torch.from_numpy(g['eps'])
    assert ((eps - ref).norm() / ref.norm()).item() < 1e-5 and tuple(taps['mid_attn'].shape[-2:]) == (16, 16)
    assert ((taps['mid_attn'] - torch.from_numpy(g['tap_mid_attn'])).norm() / torch.from_numpy(g['tap_mid_attn']).norm()).item() < 1e-5


@pytest.mark.timeout(600)
def test_full_unet_matches_reference_golden(golden_dir):
    from oracle import unet_oracle as uo
    g = np.load(f'{golden_dir}/unet_full.npz')
    sd = uo.make_params(uo.FULL, seed=0)
    x, cond = _inputs(uo.FULL, 1, 1)
    with torch.no_grad():
        eps = uo.unet_forward(sd, uo.FULL, x, uo.alpha_cosine_log_snr(torch.from_numpy(g['t'])), cond)
    ref = torch.from_numpy(g['eps'])
    assert ref.abs().max() > 0.1, 'golden eps must not be the all-zero output of a zero-initialised final_conv'
    assert ((eps - ref).norm() / ref.norm()).item() < 1e-5


def test_plms_matches_reference_trajectories(golden_dir):
    from oracle import unet_oracle as uo
    g = np.load(f'{golden_dir}/plms_small.npz')
    cfg = uo.SMALL
    sd = uo.make_params(cfg, seed=0)
    x, cond = _inputs(cfg, 1, 3)
    # 0.37: 38 calls = the expected run length of a distillation step; 0.99: the `max_thres >= .99` branch (50 steps from t = 1); None: sample() from noise
    for max_thres, expect_calls in ((0.004, 0), (0.013, 2), (0.05, 6), (0.21, 22), (0.37, 38), (0.99, 51), (None, 51)):
        key = 'noise' if max_thres is None else f'{max_thres:.3f}'
        src = uo.NoiseSource(seed=7)
        with torch.no_grad():
            start = x if max_thres is not None else src(torch.empty(x.shape))
            img, x_noisy, n0, acp, calls = uo.plms_sample(lambda xx, ls: uo.unet_forward(sd, cfg, xx, ls, cond), start, .999 if max_thres is None else max_thres, src)
        assert calls == expect_calls == int(g[f'calls_{key}'])
        ref = torch.from_numpy(g[f'img_{key}'])
        assert ((img - ref).norm() / ref.norm()).item() < 1e-4
        assert torch.allclose(x_noisy, torch.from_numpy(g[f'x_noisy_{key}']), atol=1e-6)
        assert torch.allclose(acp, torch.from_numpy(g[f'acp_{key}']), atol=1e-7)


def test_plms_call_count_law():
    """n_steps = min(int(max_thres*100), 50) -> n_steps+1 UNet calls"""
    from oracle import unet_oracle as uo
    assert [uo.plms_n_steps(t) for t in (0.0, 0.0099, 0.01, 0.10, 0.37, 0.5, 0.98)] == [0, 0, 1, 10, 37, 50, 50]
