"""diagnostic: does row i of a batched call equal the single call, for every row?  UNet (eager + graph), VAE encode / decode, PLMS sampler"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import unet_oracle as uo, vae_oracle as vo
from sparsefusion_b200.imagen_pytorch import Unet, UnetGraph
from sparsefusion_b200.ldm_autoencoder import AutoencoderKL
from sparsefusion_b200.vldm import DDPM
from sparsefusion_b200.plms import PLMSSampler
from sparsefusion_b200.distillation import KeyedNoise

dev = 'cuda'
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30)).item()
g = torch.Generator(device=dev).manual_seed(0)
for name, cfg in (('SMALL', uo.SMALL), ('FULL', uo.FULL)):
    unet = Unet(channels=cfg.channels, dim=cfg.dim, dim_mults=cfg.dim_mults, num_resnet_blocks=cfg.num_resnet_blocks, layer_attns=cfg.layer_attns,
                layer_cross_attns=(False,) * 4, cond_images_channels=cfg.cond_images_channels, attn_pool_text=False, attn_dim_head=cfg.attn_dim_head,
                attn_heads=cfg.attn_heads, cond_on_z=False, conditional_embed_dim=None)
    unet.load_state_dict(uo.make_params(cfg, seed=0))
    unet = unet.cuda()
    h = cfg.image_size
    B = 4
    x = torch.randn(B, 4, h, h, device=dev, generator=g)
    c = torch.randn(B, cfg.cond_images_channels, h, h, device=dev, generator=g)
    ls = uo.alpha_cosine_log_snr(torch.full((B,), 0.05)).cuda()
    eb = unet.forward(x, ls, cond_images=c)
    runner = UnetGraph(unet)
    gb = runner(x, ls, c).clone()
    for i in range(B):
        e1 = unet.forward(x[i:i + 1], ls[i:i + 1], cond_images=c[i:i + 1])
        print(f'{name} UNet row {i}: eager batch vs single {rel(eb[i:i+1], e1):.2e}; graph batch vs single {rel(gb[i:i+1], e1):.2e}')
    # the sampler: batch of 4 vs singles with keyed noise
    ddpm = DDPM(channels=4, unets=(unet,), conditional_encoder=None, conditional_embed_dim=None, image_sizes=(h,), timesteps=500, cond_drop_prob=0.1,
                pred_objectives='noise', conditional=False, auto_normalize_img=False, clip_output=True, dynamic_thresholding=False,
                dynamic_thresholding_percentile=.68, clip_value=10).cuda()
    lat = torch.randn(B, 4, h, h, device=dev, generator=g) * 0.2
    s = PLMSSampler(ddpm, 50)
    s.noise_fn = KeyedNoise([100 + i for i in range(B)], dev)
    pb = s.sample(lat, cond_images=c, use_tqdm=False, return_noise=True, max_thres=0.05)[0].clone()
    for i in range(B):
        s1 = PLMSSampler(ddpm, 50)
        s1.noise_fn = KeyedNoise([100 + i], dev)
        p1 = s1.sample(lat[i:i + 1], cond_images=c[i:i + 1], use_tqdm=False, return_noise=True, max_thres=0.05)[0]
        print(f'{name} PLMS row {i}: batch vs single {rel(pb[i:i+1], p1):.2e}')
    del unet, ddpm, runner
for name, kw, size in (('narrow', dict(ch=32, ch_mult=(1, 2, 4, 4)), 128), ('full', dict(), 256)):
    vae = AutoencoderKL(**kw)
    vae.load_state_dict(vo.make_params(seed=0, **kw))
    vae = vae.cuda().eval()
    B = 4
    img = torch.rand(B, 3, size, size, device=dev, generator=g) * 2 - 1
    z = torch.randn(B, 4, size // 8, size // 8, device=dev, generator=g)
    with torch.no_grad():
        mb, db = vae.encode(img).mode(), vae.decode(z)
        for i in range(B):
            print(f'{name} VAE row {i}: encode batch vs single {rel(mb[i:i+1], vae.encode(img[i:i+1]).mode()):.2e}; decode {rel(db[i:i+1], vae.decode(z[i:i+1])):.2e}')
