"""oracle/gen_golden.py -- mint the golden vectors under tests/golden/.

Run in the BUILD container only (it imports the reference from /root/reference, which does not
exist on the GPU box):

    python oracle/gen_golden.py            # all
    python oracle/gen_golden.py unet plms  # a subset

UNet / PLMS vectors are outputs of the REFERENCE's own classes (external.imagen_pytorch.Unet,
sparsefusion.vldm.DDPM, external.plms.PLMSSampler, imported unmodified) on the deterministic
weights of oracle.unet_oracle.make_params -- so they pin the restatement in oracle/unet_oracle.py
to the reference.  The reference NGP path cannot be imported here (CUDA-only extensions plus
missing trimesh/mcubes/...), so the NGP vectors are minted from oracle/ngp_oracle
itself and cross-checked against the reference CUDA sources on the GPU box
(tests/test_ref_cuda_gpu.py, oracle/_ref/).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')
REF = '/root/reference'


def _inputs(cfg, batch, seed):
    rng = np.random.default_rng(seed)
    h = cfg.image_size
    x = torch.from_numpy(rng.standard_normal((batch, cfg.channels, h, h), dtype=np.float32))
    cond = torch.from_numpy(rng.standard_normal((batch, cfg.cond_images_channels, h, h), dtype=np.float32))
    t = torch.tensor([0.37, 0.02, 0.98, 0.5][:batch], dtype=torch.float32)
    return x, cond, t


def _reference_unet(cfg, sd):
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    from external.imagen_pytorch import Unet
    unet = Unet(**cfg.reference_kwargs())
    missing = unet.load_state_dict(sd, strict=True)
    return unet.eval()


def gen_unet():
    from oracle import unet_oracle as uo
    for tag, cfg, batch in (('small', uo.SMALL, 2), ('full', uo.FULL, 1)):
        sd = uo.make_params(cfg, seed=0)
        unet = _reference_unet(cfg, sd)
        assert set(unet.state_dict().keys()) == set(sd.keys())
        x, cond, t = _inputs(cfg, batch, seed=1)
        log_snr = uo.alpha_cosine_log_snr(t)
        with torch.no_grad():
            eps = unet.forward_with_cond_scale(x, log_snr, cond_images=cond, cond_scale=1.0)
            taps = {}
            mine = uo.unet_forward(sd, cfg, x, log_snr, cond, taps)
        rel = ((mine - eps).norm() / eps.norm()).item()
        print(f'[unet {tag}] reference eps norm {eps.norm():.4f}  oracle-vs-reference rel {rel:.3e}')
        assert rel < 1e-5
        keep = {k: v.numpy() for k, v in taps.items() if k in ('init_conv', 't', 'c', 'downs.0.1', 'downs.3.3', 'mid_attn', 'ups.0.2', 'ups.0.3', 'final_res_block')}
        np.savez_compressed(os.path.join(GOLD, f'unet_{tag}.npz'), eps=eps.numpy(), t=t.numpy(), batch=batch, seed_inputs=1,
                            seed_params=0, **({f'tap_{k}': v for k, v in keep.items()} if tag == 'small' else {}))


def gen_plms():
    from oracle import unet_oracle as uo
    sys.path.insert(0, REF)
    from sparsefusion.vldm import DDPM
    from external.plms import PLMSSampler
    import external.plms as plms_mod
    cfg = uo.SMALL
    sd = uo.make_params(cfg, seed=0)
    unet = _reference_unet(cfg, sd)
    ddpm = DDPM(channels=cfg.channels, unets=(unet,), conditional_encoder=None, conditional_embed_dim=None,
                image_sizes=(cfg.image_size,), timesteps=500, cond_drop_prob=0.1, pred_objectives='noise',
                conditional=False, auto_normalize_img=False, clip_output=True, dynamic_thresholding=False,
                dynamic_thresholding_percentile=.68, clip_value=10)
    ddpm.unets[0].load_state_dict(sd)
    ddpm.eval()
    sampler = PLMSSampler(ddpm, 50)
    x, cond, _ = _inputs(cfg, 1, seed=3)
    out = {}
    for max_thres in (0.004, 0.013, 0.05, 0.21):
        src = uo.NoiseSource(seed=7)
        orig = torch.randn_like
        torch.randn_like = lambda t, **kw: src(t)  # every draw in plms.py goes through randn_like
        try:
            img, x_noisy, noise, acp = sampler.sample(x, cond_images=cond, use_tqdm=False, return_noise=True, max_thres=max_thres)
        finally:
            torch.randn_like = orig
        src2 = uo.NoiseSource(seed=7)
        with torch.no_grad():
            mine = uo.plms_sample(lambda xx, ls: uo.unet_forward(sd, cfg, xx, ls, cond), x, max_thres, src2)
        rel = ((mine[0] - img).norm() / img.norm()).item()
        print(f'[plms max_thres={max_thres}] unet calls {mine[4]} draws {src.count}/{src2.count} rel {rel:.3e}')
        assert src.count == src2.count and rel < 1e-4
        key = f'{max_thres:.3f}'
        out[f'img_{key}'] = img.numpy(); out[f'x_noisy_{key}'] = x_noisy.numpy()
        out[f'noise_{key}'] = noise.numpy(); out[f'acp_{key}'] = acp.numpy(); out[f'calls_{key}'] = mine[4]
    np.savez_compressed(os.path.join(GOLD, 'plms_small.npz'), **out)


def gen_vae():
    """the plain-torch VAE mirror vs the reference's Encoder / Decoder (external/ldm/modules/diffusionmodules/model.py)"""
    sys.path.insert(0, REF)
    from external.ldm.modules.diffusionmodules.model import Encoder as REnc, Decoder as RDec
    from sparsefusion_b200.ldm_autoencoder import AutoencoderKL
    dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2,
              attn_resolutions=[], dropout=0.0)
    torch.manual_seed(0)
    renc, rdec = REnc(**dd).eval(), RDec(**dd).eval()
    vae = AutoencoderKL().eval()
    vae.encoder.load_state_dict(renc.state_dict(), strict=True)
    vae.decoder.load_state_dict(rdec.state_dict(), strict=True)
    x, z = torch.randn(1, 3, 64, 64), torch.randn(1, 4, 8, 8)
    with torch.no_grad():
        de, dd_ = (renc(x) - vae.encoder(x)).abs().max().item(), (rdec(z) - vae.decoder(z)).abs().max().item()
    print(f'[vae] mirror vs reference Encoder/Decoder: max abs diff {de:.1e} / {dd_:.1e}')
    assert de == 0.0 and dd_ == 0.0


def gen_ngp():
    from oracle import ngp_oracle as no
    no.write_golden(GOLD)


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ['unet', 'plms', 'ngp', 'vae']
    torch.set_num_threads(os.cpu_count())
    if 'unet' in which:
        gen_unet()
    if 'plms' in which:
        gen_plms()
    if 'ngp' in which:
        gen_ngp()
    if 'vae' in which:
        gen_vae()
