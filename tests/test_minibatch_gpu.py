"""The view-batched step (Distiller.minibatch_step): one optimiser step on grad(photometric) + mean over V target views.

(1) composition: its gradient equals grad(A) + (1/V) sum_v grad(B_v) with every term computed by the one-view sub-steps (the reference-shaped
    code path that tests/test_distillation_gpu.py pins to the restatement), same injected draws;
(2) R-independence: the same V-view step run on 1 rank and sharded over 2 ranks (gloo over CUDA tensors) yields the same gradient and the same
    parameters -- every per-view draw comes from a generator keyed by (seed, iteration, view), so nothing depends on which rank or batch a view
    lands in -- with ONE all-reduce per step;
(3) batching: the PLMS / VAE outputs of a view do not depend on the batch it shares (UNet batch 1 vs 2 vs 4).
"""
import os
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _build(rank, world, device, views_per_step, percep=False):
    from oracle import distill_oracle as do, ngp_oracle as no, unet_oracle as uo, vae_oracle as vo
    from sparsefusion_b200.distillation import Distiller, SceneCache
    from sparsefusion_b200.imagen_pytorch import Unet
    from sparsefusion_b200.ldm_autoencoder import AutoencoderKL
    from sparsefusion_b200.network_grid import NeRFNetwork, get_default_torch_ngp_opt
    from sparsefusion_b200.vldm import DDPM
    cfg = uo.SMALL
    scene = do.synthetic_scene(n_input=2, n_target=6, image_size=128, latent=16, feat_ch=cfg.cond_images_channels, render_hw=64, seed=3)
    vae = AutoencoderKL(ch=32, ch_mult=(1, 2, 4, 4)).eval()
    vae.load_state_dict(vo.make_params(seed=0, ch=32, ch_mult=(1, 2, 4, 4)))
    opt = get_default_torch_ngp_opt()
    ngp = NeRFNetwork(opt)
    st = ngp.state_dict()
    st.update(no.make_field_params(seed=0))
    ngp.load_state_dict(st)
    ngp = ngp.to(device).train()
    unet = Unet(channels=cfg.channels, dim=cfg.dim, dim_mults=cfg.dim_mults, num_resnet_blocks=cfg.num_resnet_blocks, layer_attns=cfg.layer_attns,
                layer_cross_attns=(False,) * 4, cond_images_channels=cfg.cond_images_channels, attn_pool_text=False,
                attn_dim_head=cfg.attn_dim_head, attn_heads=cfg.attn_heads, cond_on_z=False, conditional_embed_dim=None)
    unet.load_state_dict(uo.make_params(cfg, seed=0))
    ddpm = DDPM(channels=4, unets=(unet.to(device),), conditional_encoder=None, conditional_embed_dim=None, image_sizes=(cfg.image_size,), timesteps=500,
                cond_drop_prob=0.1, pred_objectives='noise', conditional=False, auto_normalize_img=False, clip_output=True,
                dynamic_thresholding=False, dynamic_thresholding_percentile=.68, clip_value=10).to(device)
    pl = None
    if percep:
        from sparsefusion_b200.lpips_vgg import PerceptualLoss
        pl = PerceptualLoss('vgg', device=device, seed=2)
    return Distiller(ngp, vae.to(device), ddpm, opt, SceneCache(**scene).to(device), seed=11, rank=rank, world_size=world, process_group=None,
                     views_per_step=views_per_step, percep=pl)


def _fixed_target(views, pred_img):
    """deterministic stand-in for the decoded PLMS sample (see tests/test_multirank_gpu.py): one image per VIEW, whatever batch it arrives in"""
    out = []
    for v in views:
        g = torch.Generator().manual_seed(1000 + v)
        out.append(torch.rand(pred_img.shape[1:], generator=g))
    return torch.stack(out).to(pred_img.device)


def _no_update(dist):
    dist.optimizer.step = lambda grad_scale=1.0: None


@pytest.mark.parametrize('itr', [5, 1500])
def test_minibatch_gradient_is_photometric_plus_mean_of_view_gradients(itr):
    from sparsefusion_b200.distillation import step_views, view_seed
    dev = torch.device('cuda', 0)
    V = 3
    mb = _build(0, 1, dev, V, percep=(itr > 1000))
    mb.pred_img_hook = _fixed_target
    _no_update(mb)
    gen_state = mb.gen.get_state()
    la, lb = mb.minibatch_step(itr, max_thres=0.05)
    g_mb = mb.optimizer.grad.clone()
    views, calls = mb.last['views'], mb.last.get('unet_calls')
    assert len(views) == V and len(set(views)) == V
    assert calls == (6 if itr > 1000 else None)
    # the same draws the keyed generators made, replayed through the one-view sub-steps
    g = torch.Generator().manual_seed(11)
    g.set_state(gen_state)
    idx = int(torch.randperm(2, generator=g)[0])
    assert step_views(torch.randperm(6, generator=g), V) == views

    def keyed_noise(n, seed):
        gg = torch.Generator(device=dev)
        gg.manual_seed(seed)
        return torch.rand(n, 64, generator=gg, device=dev), torch.rand(n, 64, generator=gg, device=dev)
    N = 64 * 64
    single = _build(0, 1, dev, None, percep=(itr > 1000))
    single.pred_img_hook = _fixed_target
    _no_update(single)
    # A: photometric_substep draws its own input-view index from single.gen: force the same one
    single.gen.set_state(gen_state)
    single.render_noise = lambda k: keyed_noise(N, view_seed(11, itr, idx, 0))
    l_a = single.photometric_substep(itr)
    want = single.optimizer.grad.clone()
    l_b = 0.0
    for v in views:
        single.render_noise = lambda k, v=v: keyed_noise(N, view_seed(11, itr, v, 1))
        # fusion_substep picks perm[1 + rank]: present view v as that entry by running the fused sub-step directly
        single.optimizer.zero_grad()
        u = torch.zeros(1)
        from sparsefusion_b200.distillation import KeyedNoise
        single.sampler.noise_fn = KeyedNoise([view_seed(11, itr, v, 2)], dev)
        lv = single._fusion_substep_fused(itr, 0.05, v, u, single.cache.target_features[v:v + 1])
        want += single.optimizer.grad / V
        l_b = l_b + lv / V
    rel = ((g_mb - want).norm() / want.norm()).item()
    print(f'itr {itr}: minibatch gradient vs grad(A) + mean_v grad(B_v): rel {rel:.3e}; losses {la.item():.6f}/{l_a.item():.6f}  {lb.item():.6f}/{float(l_b):.6f}')
    assert rel < 1e-4
    assert abs(la.item() - l_a.item()) < 1e-5 * abs(l_a.item()) and abs(lb.item() - float(l_b)) < 1e-4 * abs(float(l_b))


def _worker(rank, world, store, out_dir, itr, V):
    import torch.distributed as td
    td.init_process_group('gloo', init_method=f'file://{store}', rank=rank, world_size=world)
    try:
        device = torch.device('cuda', rank % torch.cuda.device_count())
        torch.cuda.set_device(device)
        dist = _build(rank, world, device, V)
        seen = {}

        def hook(views, pred_img):
            for j, v in enumerate(views):
                seen[v] = pred_img[j].cpu()
            return _fixed_target(views, pred_img)
        dist.pred_img_hook = hook
        n_coll = [0]
        orig = td.all_reduce

        def counting(*a, **k):
            n_coll[0] += 1
            return orig(*a, **k)
        td.all_reduce = counting
        dist.minibatch_step(itr, max_thres=0.05)
        td.all_reduce = orig
        torch.cuda.synchronize()
        torch.save(dict(grad=dist.optimizer.grad.cpu(), flat=dist.optimizer.flat.cpu(), views=dist.last['views'], seen=seen, collectives=n_coll[0]),
                   os.path.join(out_dir, f'r{rank}.pt'))
    finally:
        td.destroy_process_group()


@pytest.mark.timeout(900)
@pytest.mark.parametrize('itr', [5, 1500])
def test_minibatch_step_is_independent_of_the_number_of_ranks(itr):
    import torch.multiprocessing as mp
    V, world = 4, 2
    with tempfile.TemporaryDirectory() as d:
        try:
            mp.spawn(_worker, args=(world, os.path.join(d, 'store'), d, itr, V), nprocs=world, join=True)
        except Exception as e:   # noqa: BLE001
            if 'gloo' in str(e).lower() and 'cuda' in str(e).lower():
                pytest.skip(f'gloo without CUDA tensor support in this torch build: {e}')
            raise
        r = [torch.load(os.path.join(d, f'r{k}.pt')) for k in range(world)]
    assert r[0]['collectives'] == r[1]['collectives'] == 1                       # ONE all-reduce per step
    assert torch.equal(r[0]['grad'], r[1]['grad']) and torch.equal(r[0]['flat'], r[1]['flat'])     # replicas stay bit-identical
    assert len(r[0]['views']) == len(r[1]['views']) == V // world and not set(r[0]['views']) & set(r[1]['views'])
    dev = torch.device('cuda', 0)
    one = _build(0, 1, dev, V)
    seen = {}

    def hook(views, pred_img):
        for j, v in enumerate(views):
            seen[v] = pred_img[j].cpu()
        return _fixed_target(views, pred_img)
    one.pred_img_hook = hook
    flat0 = one.optimizer.flat.clone().cpu()
    one.minibatch_step(itr, max_thres=0.05)
    assert sorted(one.last['views']) == sorted(r[0]['views'] + r[1]['views'])
    rel_g = ((r[0]['grad'] - one.optimizer.grad.cpu()).norm() / one.optimizer.grad.cpu().norm()).item()
    upd = one.optimizer.flat.cpu() - flat0
    rel_p = ((r[0]['flat'] - one.optimizer.flat.cpu()).norm() / upd.norm()).item()
    print(f'itr {itr}: 2-rank vs 1-rank {V}-view step: gradient rel {rel_g:.3e}, parameter update rel {rel_p:.3e}')
    assert rel_g < 1e-4 and rel_p < 1e-2
    if itr > 1000:      # the denoised image of a view must not depend on the batch it was sampled in (UNet / VAE batch 2 on each rank vs 4 here)
        both = {**r[0]['seen'], **r[1]['seen']}
        for v, img in seen.items():
            rel = ((both[v] - img).norm() / img.norm()).item()
            assert rel < 1e-3, (v, rel)
