"""uninitialised-read detector: fill the caching allocator's free blocks with NaN, run an operator, look for NaN / run-to-run differences"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as uo, vae_oracle as vo
from sparsefusion_b200.imagen_pytorch import Unet
from sparsefusion_b200.ldm_autoencoder import AutoencoderKL

dev = 'cuda'


def poison():
    """leave NaN in every size class the operators are likely to reuse"""
    blocks = []
    for mb in (0.001, 0.004, 0.016, 0.06, 0.25, 1, 2, 4, 8, 16, 32, 64, 128, 256):
        for _ in range(6):
            blocks.append(torch.full((max(1, int(mb * (1 << 20) // 4)),), float('nan'), device=dev))
    torch.cuda.synchronize()
    del blocks


def check(name, fn):
    torch.cuda.empty_cache()
    ref = fn().clone()
    bad = 0
    worst = 0.0
    for i in range(4):
        poison()
        out = fn()
        torch.cuda.synchronize()
        if not torch.isfinite(out).all():
            bad += 1
        else:
            worst = max(worst, ((out.double() - ref.double()).norm() / ref.double().norm()).item())
    print(f'{name:46s} non-finite runs {bad}/4   worst rel vs clean run {worst:.2e}', flush=True)


g = torch.Generator(device=dev).manual_seed(0)
for name, kw, size in (('narrow', dict(ch=32, ch_mult=(1, 2, 4, 4)), 128), ('full', dict(), 256)):
    vae = AutoencoderKL(**kw)
    vae.load_state_dict(vo.make_params(seed=0, **kw))
    vae = vae.cuda().eval()
    for nb in (1, 2):
        img = torch.rand(nb, 3, size, size, device=dev, generator=g) * 2 - 1
        z = torch.randn(nb, 4, size // 8, size // 8, device=dev, generator=g)
        with torch.no_grad():
            check(f'VAE {name} encode batch {nb}', lambda: vae.encode(img).mode())
            check(f'VAE {name} decode batch {nb}', lambda: vae.decode(z))
    # stage by stage for the decode at batch 1
    eng = vae._sm100
    from sparsefusion_b200 import ops
    with torch.no_grad():
        z = torch.randn(1, 4, size // 8, size // 8, device=dev, generator=g)
        zin = torch.zeros(1, size // 8, size // 8, 4, device=dev)
        ops.nchw_to_nhwc(z, zin, 0)
        h0 = eng.conv('decoder.conv_in', eng.conv('post_quant_conv', zin, 1), 3).clone()
        check(f'  {name} conv_in', lambda: eng.conv('decoder.conv_in', eng.conv('post_quant_conv', zin, 1), 3))
        check(f'  {name} mid.block_1', lambda: eng.resnet('decoder.mid.block_1', h0))
        check(f'  {name} mid.attn_1', lambda: eng.attn('decoder.mid.attn_1', h0))
        check(f'  {name} groupnorm', lambda: eng.gn('decoder.mid.block_1.norm1', h0, True))
        h1 = eng.resnet('decoder.mid.block_1', h0).clone()
        check(f'  {name} up.3.block.0', lambda: eng.resnet('decoder.up.3.block.0', h1))
        up = ops.upsample2x(h1).clone()
        check(f'  {name} upsample conv', lambda: eng.conv('decoder.up.3.upsample.conv', up, 3))
for name, cfg in (('SMALL', uo.SMALL), ('FULL', uo.FULL)):
    unet = Unet(channels=cfg.channels, dim=cfg.dim, dim_mults=cfg.dim_mults, num_resnet_blocks=cfg.num_resnet_blocks, layer_attns=cfg.layer_attns,
                layer_cross_attns=(False,) * 4, cond_images_channels=cfg.cond_images_channels, attn_pool_text=False, attn_dim_head=cfg.attn_dim_head,
                attn_heads=cfg.attn_heads, cond_on_z=False, conditional_embed_dim=None)
    unet.load_state_dict(uo.make_params(cfg, seed=0))
    unet = unet.cuda()
    h = cfg.image_size
    for nb in (1, 2):
        x = torch.randn(nb, 4, h, h, device=dev, generator=g)
        c = torch.randn(nb, cfg.cond_images_channels, h, h, device=dev, generator=g)
        ls = uo.alpha_cosine_log_snr(torch.full((nb,), 0.05)).cuda()
        check(f'UNet {name} eager batch {nb}', lambda: unet.forward(x, ls, cond_images=c))
from sparsefusion_b200.lpips_vgg import PerceptualLoss
pl = PerceptualLoss('vgg', device=dev, seed=0)
a, b = torch.rand(3, 128, 128, device=dev, generator=g), torch.rand(3, 128, 128, device=dev, generator=g)
check('LPIPS value', lambda: pl.value_and_grad(a, b)[0].reshape(1))
check('LPIPS gradient', lambda: pl.value_and_grad(a, b)[1])
