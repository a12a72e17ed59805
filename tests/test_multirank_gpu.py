"""the view-sharded fusion sub-step on R ranks must equal the single-rank R-view minibatch step.

Two processes (gloo over CUDA tensors -- one GPU is enough; NCCL needs one device per rank and is exercised by bench.py --gpus 2) run
`Distiller.fusion_substep` with world_size 2; the parent recomputes the two views' gradients with world_size 1 and checks that the all-reduced
flat gradient is their sum, that both ranks end with bit-identical parameters, and that those equal a single Adam step on the mean gradient.
"""
import os
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(rank, world, pg, device):
    from oracle import distill_oracle as do, ngp_oracle as no, unet_oracle as uo
    from sparsefusion_b200.distillation import Distiller, SceneCache
    from sparsefusion_b200.imagen_pytorch import Unet
    from sparsefusion_b200.ldm_autoencoder import AutoencoderKL
    from sparsefusion_b200.network_grid import NeRFNetwork, get_default_torch_ngp_opt
    from sparsefusion_b200.vldm import DDPM
    cfg = uo.SMALL
    scene = do.synthetic_scene(n_input=2, n_target=6, image_size=128, latent=16, feat_ch=cfg.cond_images_channels, render_hw=64, seed=3)
    torch.manual_seed(0)
    vae = AutoencoderKL(ch=32, ch_mult=(1, 2, 4, 4)).eval()
    from oracle import vae_oracle as vo
    ref_vae = vo.TorchVAE(vae.state_dict())          # the checker's VAE: torch restatement over the same weights
    sd = uo.make_params(cfg, seed=0)
    p = no.make_field_params(seed=0)
    opt = get_default_torch_ngp_opt()
    ngp = NeRFNetwork(opt)
    st = ngp.state_dict()
    st.update({k: v for k, v in p.items()})
    ngp.load_state_dict(st)
    ngp = ngp.to(device).train()
    unet = Unet(channels=cfg.channels, dim=cfg.dim, dim_mults=cfg.dim_mults, num_resnet_blocks=cfg.num_resnet_blocks, layer_attns=cfg.layer_attns,
                layer_cross_attns=(False,) * 4, cond_images_channels=cfg.cond_images_channels, attn_pool_text=False,
                attn_dim_head=cfg.attn_dim_head, attn_heads=cfg.attn_heads, cond_on_z=False, conditional_embed_dim=None)
    unet.load_state_dict(sd)
    ddpm = DDPM(channels=4, unets=(unet.to(device),), conditional_encoder=None, conditional_embed_dim=None, image_sizes=(cfg.image_size,), timesteps=500,
                cond_drop_prob=0.1, pred_objectives='noise', conditional=False, auto_normalize_img=False, clip_output=True,
                dynamic_thresholding=False, dynamic_thresholding_percentile=.68, clip_value=10).to(device)
    dist = Distiller(ngp, vae.to(device), ddpm, opt, SceneCache(**scene).to(device), seed=11, rank=rank, world_size=world, process_group=pg)
    N = 64 * 64
    rng = np.random.default_rng(100 + rank)                          # per-rank render noise, reproducible in the single-rank reference
    noise = (torch.from_numpy(rng.random((N, 64), dtype=np.float32)).to(device), torch.from_numpy(rng.random((N, 64), dtype=np.float32)).to(device))
    dist.render_noise = lambda k: noise
    src = uo.NoiseSource(seed=50 + rank)
    dist.sampler.noise_fn = lambda t: src(t.cpu()).to(t.device)
    # SDS phase: the loss is an L1 against the decoded PLMS sample, and that sample differs at 1e-6 between any two runs (split-K reduction order
    # of the UNet), which flips sign(image - pred) on near-tie pixels.  The denoised target is therefore replaced on both sides by a fixed image per
    # view (the sampler and the VAE still run; their output is checked in tests/test_distillation_gpu.py): what is compared here is the sharding,
    # the collective and the update, deterministically.
    def fixed_target(views, pred_img):
        g = torch.Generator().manual_seed(1000 + views[0])
        return torch.rand(pred_img.shape, generator=g).to(pred_img.device)
    dist.pred_img_hook = fixed_target
    return dist


def _worker(rank, world, store, out_dir, itr):
    import torch.distributed as td
    td.init_process_group('gloo', init_method=f'file://{store}', rank=rank, world_size=world)
    try:
        device = torch.device('cuda', rank % torch.cuda.device_count())
        torch.cuda.set_device(device)
        dist = _build(rank, world, None, device)
        dist.fusion_substep(itr, max_thres=0.05)
        torch.cuda.synchronize()
        torch.save(dict(grad=dist.optimizer.grad.cpu(), flat=dist.optimizer.flat.cpu(), calls=dist.last.get('unet_calls')), os.path.join(out_dir, f'r{rank}.pt'))
    finally:
        td.destroy_process_group()


# itr 5: EFT-bootstrap loss (smooth huber, no UNet); itr 1500: SDS loss with the denoised target injected on both sides (see _build) -> in both
# cases everything but the float atomics of the grid backward is deterministic: tight bounds.
@pytest.mark.timeout(600)
@pytest.mark.parametrize('itr,tol', [(5, 1e-4), (1500, 1e-3)])
def test_two_rank_fusion_step_equals_two_view_minibatch(itr, tol):
    import torch.multiprocessing as mp
    world = 2
    with tempfile.TemporaryDirectory() as d:
        try:
            mp.spawn(_worker, args=(world, os.path.join(d, 'store'), d, itr), nprocs=world, join=True)
        except Exception as e:   # noqa: BLE001
            if 'gloo' in str(e).lower() and 'cuda' in str(e).lower():
                pytest.skip(f'gloo without CUDA tensor support in this torch build: {e}')
            raise
        r = [torch.load(os.path.join(d, f'r{k}.pt')) for k in range(world)]
    assert torch.equal(r[0]['grad'], r[1]['grad']) and torch.equal(r[0]['flat'], r[1]['flat'])       # ranks stay bit-identical
    assert r[0]['calls'] == r[1]['calls'] == (6 if itr > 1000 else None)
    # single-rank reference: the same two views, one after the other, gradients summed by hand
    dev = torch.device('cuda', 0)
    grads = []
    for k in range(world):
        single = _build(k, 1, None, dev)            # rank k's view / noise, but no collective (world_size 1)
        flat0 = single.optimizer.flat.clone()
        orig_step = single.optimizer.step
        single.optimizer.step = lambda grad_scale=1.0: None          # keep the parameters: only the gradient of this view is wanted
        single.fusion_substep(itr, max_thres=0.05)
        grads.append(single.optimizer.grad.clone())
        single.optimizer.step = orig_step
    want = (grads[0] + grads[1]).cpu()
    rel = ((r[0]['grad'] - want).norm() / want.norm()).item()
    print(f'all-reduced gradient vs sum of the two single-rank gradients: rel {rel:.3e}')
    assert rel < tol
    # and the update is ONE Adam step on the mean gradient
    single.optimizer.flat.copy_(flat0)
    single.optimizer.grad.copy_(want.to(dev))
    single.optimizer.m.zero_(); single.optimizer.v.zero_(); single.optimizer.t = 0
    single.optimizer.step(grad_scale=0.5)
    relp = ((r[0]['flat'] - single.optimizer.flat.cpu()).norm() / (single.optimizer.flat.cpu() - flat0.cpu()).norm()).item()
    print(f'parameters after the 2-rank step vs one Adam step on the mean gradient: rel (of the update) {relp:.3e}')
    assert relp < max(1e-2, 10 * tol)
