"""oracle/vae_oracle.py (torch restatement of the SD KL-f8 autoencoder) against tests/golden/vae.npz, which holds outputs of the REFERENCE's own
Encoder / Decoder modules (external/ldm/modules/diffusionmodules/model.py:368-568) minted by oracle/gen_golden.py `vae`.  Same torch ops in the
same order -> bit-identical on the same torch build; a tolerance of 1e-6 keeps the test meaningful across torch versions."""
import numpy as np
import torch


def test_restatement_reproduces_reference_modules(golden_dir):
    from oracle import vae_oracle as vo
    g = np.load(f'{golden_dir}/vae.npz')
    for tag, cfg in (('full', dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2)), ('narrow', dict(ch=32, ch_mult=(1, 2, 4, 4), num_res_blocks=2))):
        sd = vo.make_params(seed=int(g['param_seed']), **cfg)
        size = int(g[f'{tag}_size'])
        rng = np.random.default_rng(int(g['input_seed']))
        x = torch.from_numpy(rng.random((1, 3, size, size), dtype=np.float32) * 2 - 1)
        z = torch.from_numpy(rng.standard_normal((1, 4, size // 8, size // 8), dtype=np.float32))
        vae = vo.TorchVAE(sd)
        post = vae.encode(x)
        mom = torch.from_numpy(g[f'{tag}_moments'])
        assert torch.allclose(post.mean, mom[:, :4], atol=1e-6) and torch.allclose(post.logvar, mom[:, 4:].clamp(-30, 20), atol=1e-6)
        assert torch.allclose(vae.decode(z), torch.from_numpy(g[f'{tag}_dec']), atol=1e-6)


def test_state_dict_layout_is_the_products_and_the_product_has_no_torch_path():
    """the product's AutoencoderKL is a parameter container with the same keys / shapes; calling it on CPU raises"""
    import pytest
    from oracle import vae_oracle as vo
    from sparsefusion_b200.ldm_autoencoder import AutoencoderKL
    vae = AutoencoderKL()
    sh = vo.param_shapes()
    sd = vae.state_dict()
    assert set(sh) == set(sd) and all(tuple(sd[k].shape) == sh[k] for k in sh)
    vae.load_state_dict(vo.make_params(seed=3), strict=True)
    with pytest.raises(RuntimeError), torch.no_grad():
        vae.encode(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError), torch.no_grad():
        vae.decoder(torch.zeros(1, 4, 8, 8))
