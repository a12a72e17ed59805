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


def gen_unet(only=()):
    from oracle import unet_oracle as uo
    import dataclasses
    # small128: BASELINE configs[4] geometry (128x128x4 latents: 16x16 = 256 tokens + 2 time tokens + null key at the attention stage, 16 384-pixel
    # convolutions) at the narrow width, so that the reference itself can mint it on CPU
    which = [('small', uo.SMALL, 2), ('full', uo.FULL, 1), ('small128', dataclasses.replace(uo.SMALL, image_size=128), 1)]
    if only:
        which = [w for w in which if w[0] in only]
    for tag, cfg, batch in which:
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
                            seed_params=0, **({f'tap_{k}': v for k, v in keep.items()} if tag == 'small' else
                                              {f'tap_{k}': v for k, v in keep.items() if k in ('mid_attn', 'downs.3.3')} if tag == 'small128' else {}))


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
    # 0.37 -> 38 UNet calls (the expected length of a distillation step's run); 0.99 -> the `max_thres >= .99` branch (:80-85): 50 steps from t = 1.0,
    # where the log-SNR must be evaluated in fp32 like the reference (cos(pi/2) is -4.4e-8 there, 6e-17 in fp64); None -> sample() from noise
    for max_thres in (0.004, 0.013, 0.05, 0.21, 0.37, 0.99, None):
        src = uo.NoiseSource(seed=7)
        orig, orig_randn = torch.randn_like, torch.randn
        torch.randn_like = lambda t, **kw: src(t)  # every draw in plms.py goes through randn_like ...
        torch.randn = lambda shape, **kw: src(torch.empty(tuple(shape)))   # ... except the start image of sample() from noise (plms.py:73)
        try:
            if max_thres is None:
                img, x_noisy, noise, acp = sampler.sample(cond_images=cond, use_tqdm=False, return_noise=True)
            else:
                img, x_noisy, noise, acp = sampler.sample(x, cond_images=cond, use_tqdm=False, return_noise=True, max_thres=max_thres)
        finally:
            torch.randn_like, torch.randn = orig, orig_randn
        src2 = uo.NoiseSource(seed=7)
        with torch.no_grad():
            start = x if max_thres is not None else src2(torch.empty(x.shape))
            mine = uo.plms_sample(lambda xx, ls: uo.unet_forward(sd, cfg, xx, ls, cond), start, .999 if max_thres is None else max_thres, src2)
        rel = ((mine[0] - img).norm() / img.norm()).item()
        print(f'[plms max_thres={max_thres}] unet calls {mine[4]} draws {src.count}/{src2.count} rel {rel:.3e}')
        assert src.count == src2.count and rel < 1e-4
        key = 'noise' if max_thres is None else f'{max_thres:.3f}'
        out[f'img_{key}'] = img.numpy(); out[f'x_noisy_{key}'] = x_noisy.numpy()
        out[f'noise_{key}'] = noise.numpy(); out[f'acp_{key}'] = acp.numpy(); out[f'calls_{key}'] = mine[4]
    np.savez_compressed(os.path.join(GOLD, 'plms_small.npz'), **out)


def gen_vae():
    """tests/golden/vae.npz: outputs of the REFERENCE's own Encoder / Decoder (external/ldm/modules/diffusionmodules/model.py:368-568, imported
    unmodified) on oracle.vae_oracle.make_params weights, for the loop's full-width configuration at 64x64 and a narrow one at 128x128; the
    restatement oracle/vae_oracle.py must reproduce them bit for bit (same torch ops in the same order)."""
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    from external.ldm.modules.diffusionmodules.model import Encoder as REnc, Decoder as RDec
    from oracle import vae_oracle as vo
    out = {}
    for tag, cfg, size in (('full', dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2), 64), ('narrow', dict(ch=32, ch_mult=(1, 2, 4, 4), num_res_blocks=2), 128)):
        dd = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, attn_resolutions=[], dropout=0.0, **cfg)
        renc, rdec = REnc(**dd).eval(), RDec(**dd).eval()
        sd = vo.make_params(seed=0, **cfg)
        renc.load_state_dict({k[len('encoder.'):]: v for k, v in sd.items() if k.startswith('encoder.')}, strict=True)
        rdec.load_state_dict({k[len('decoder.'):]: v for k, v in sd.items() if k.startswith('decoder.')}, strict=True)
        rng = np.random.default_rng(17)
        x = torch.from_numpy(rng.random((1, 3, size, size), dtype=np.float32) * 2 - 1)
        z = torch.from_numpy(rng.standard_normal((1, 4, size // 8, size // 8), dtype=np.float32))
        with torch.no_grad():
            # autoencoder.py:311-323: moments = quant_conv(encoder(x)); dec = decoder(post_quant_conv(z))
            moments = torch.nn.functional.conv2d(renc(x), sd['quant_conv.weight'], sd['quant_conv.bias'])
            dec = rdec(torch.nn.functional.conv2d(z, sd['post_quant_conv.weight'], sd['post_quant_conv.bias']))
            mine = vo.TorchVAE(sd)
            post = mine.encode(x)
            de = (torch.cat((post.mean, post.logvar), 1) - torch.cat((moments[:, :4], moments[:, 4:].clamp(-30, 20)), 1)).abs().max().item()
            dd_ = (mine.decode(z) - dec).abs().max().item()
        print(f'[vae {tag}] restatement vs reference Encoder/Decoder: max abs diff {de:.1e} / {dd_:.1e}')
        assert de == 0.0 and dd_ == 0.0
        out[f'{tag}_moments'], out[f'{tag}_dec'], out[f'{tag}_size'] = moments.numpy(), dec.numpy(), size
    np.savez_compressed(os.path.join(GOLD, 'vae.npz'), input_seed=17, param_seed=0, **out)


def gen_ngp():
    from oracle import ngp_oracle as no
    no.write_golden(GOLD)


def gen_eft():
    """tests/golden/eft.npz: outputs of the REFERENCE's own EpipolarFeatureTransformer (sparsefusion/eft.py, imported unmodified; pytorch3d's RayBundle and
    the two camera methods it calls supplied by oracle/eft_oracle.py) on deterministic weights and inputs: the per-ray colour, the 256-d feature the
    VLDM is conditioned on, and samples of the ResNet-18 pyramid"""
    from oracle import eft_oracle as eo
    from sparsefusion_b200.eft import EpipolarFeatureTransformer as Mirror
    eft = eo.import_reference_eft(REF)
    try:
        ref = eft.EpipolarFeatureTransformer(use_r=True, encoder='resnet18', return_features=True, remove_unused_layers=False).eval()   # utils/load_model.py:34
    finally:
        eft._restore_resnet18()
    mirror = Mirror(use_r=True, encoder='resnet18', return_features=True, remove_unused_layers=False)
    shapes = {k: tuple(v.shape) for k, v in mirror.state_dict().items()}
    assert shapes == {k: tuple(v.shape) for k, v in ref.state_dict().items()}, 'the mirror must have the reference state_dict layout'
    ref.load_state_dict(eo.make_params(shapes, seed=0), strict=True)
    images, cams, rb = eo.scene_inputs()
    with torch.no_grad():
        _, latent = ref.encode(cams, images)
        rgb, f3, _ = ref.forward(rb)
        rgb_b, f3_b, _ = ref.batched_forward(eo.RayBundle(rb.origins.view(6, 8, 3), rb.directions.view(6, 8, 3), rb.lengths.view(6, 8, -1), None), n_batches=4)
    assert torch.allclose(rgb_b.reshape(-1, 3), rgb, atol=1e-6) and torch.allclose(f3_b.reshape(-1, 256), f3, atol=1e-5)    # chunking changes nothing
    print(f'[eft] latent {tuple(latent.shape)} norm {latent.norm():.3f}; rgb mean {rgb.mean():.4f}; f3 norm {f3.norm():.3f}; {len(shapes)} tensors')
    flat = latent.permute(0, 2, 3, 1).reshape(-1)
    np.savez_compressed(os.path.join(GOLD, 'eft.npz'), rgb=rgb.numpy(), f3=f3.numpy(), latent_shape=np.array(latent.shape), latent_norm=float(latent.norm()),
                        latent_sample=flat[::997].numpy(), n_tensors=len(shapes), param_seed=0, input_seed=3)


def _import_reference_ngp():
    """import the REFERENCE's external.nerf.network_grid (NeRFNetwork over renderer_df.NeRFRenderer) on this CPU-only container.  Its module-level
    imports that cannot be satisfied here are replaced by inert stubs (trimesh, mcubes, torch_ema, lpips, ... -- none is touched by run()), and the
    two CUDA-only operators it reaches on the default path get CPU stand-ins backed by the C restatement (which tests/test_ref_cuda_gpu.py pins
    to the reference's CUDA sources on the GPU box): `raymarching.near_far_from_aabb` (renderer_df.py:328) and the grid encoder behind
    `external.ngp_encoder.get_encoder` (network_grid.py:50).  Everything else -- sample_pdf, run(), MLP, trunc_exp, the density blob,
    common_forward / density / forward -- is the reference's own code, executed unmodified."""
    import types
    from oracle import ngp_oracle as no
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True

    class _Any:
        def __init__(self, *a, **k): pass
        def __call__(self, *a, **k): return self
        def __getattr__(self, k): return _Any()

    def stub(name):
        m = types.ModuleType(name)
        m.__getattr__ = lambda k: _Any()
        m.__path__ = []
        sys.modules[name] = m
        return m

    rm = stub('raymarching')

    def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
        n, f = no.near_far_from_aabb(rays_o.detach().numpy(), rays_d.detach().numpy(), aabb.numpy(), min_near)
        return torch.from_numpy(n), torch.from_numpy(f)
    rm.near_far_from_aabb = near_far_from_aabb
    for _ in range(40):
        try:
            import external.nerf.network_grid as ng
            break
        except ModuleNotFoundError as e:
            stub(e.name)
    import external.ngp_encoder as ne

    class CpuGridEncoder(torch.nn.Module):
        """external/gridencoder/grid.py:91-154 with the CUDA op replaced by the C restatement"""

        def __init__(self, bound):
            super().__init__()
            self.geo = no.live_geometry(bound)
            self.embeddings = torch.nn.Parameter(torch.zeros(int(self.geo['offsets'][-1]), self.geo['C']))
            self.output_dim = self.geo['L'] * self.geo['C']

        def forward(self, inputs, bound=1):
            inputs = (inputs + bound) / (2 * bound)                                  # grid.py:142
            return no._GridEncode.apply(inputs.reshape(-1, 3).contiguous(), self.embeddings, self.geo, None)

    def get_encoder(encoding, input_dim=3, **kw):
        assert encoding == 'tiledgrid' and input_dim == 3
        enc = CpuGridEncoder(kw['desired_resolution'] / 2048)
        return enc, enc.output_dim
    ng.get_encoder = get_encoder
    import external.nerf.renderer_df as rdf
    return ng, rdf


def gen_run_ref():
    """tests/golden/ngp_run_ref.npz: outputs and parameter gradients of the REFERENCE's own NeRFNetwork.run / sample_pdf on the inputs of
    ngp_run.npz -- pins the oracle's single-evaluation restatement of run() (oracle/ngp_oracle.py:338-398) to the reference's literal three
    passes (renderer_df.py:373,398,424)."""
    from types import SimpleNamespace
    from oracle import ngp_oracle as no
    ng, rdf = _import_reference_ngp()
    opt = SimpleNamespace(cuda_ray=False, max_steps=256, num_steps=64, upsample_steps=64, update_extra_interval=16, max_ray_batch=4096,
                          albedo_iters=1000, bg_radius=0, density_thresh=10, fp16=True, backbone='grid', w=128, h=128, hw_scale=2, bound=4,
                          min_near=0.1, dt_gamma=0, lambda_entropy=1e-4, lambda_opacity=0, lambda_orient=1e-2, lambda_smooth=0)   # distillation.py:500-526
    net = ng.NeRFNetwork(opt).train()
    p = no.make_field_params(seed=0)
    sd = net.state_dict()
    sd.update(p)
    net.load_state_dict(sd)
    g = np.load(os.path.join(GOLD, 'ngp_run.npz'))
    ro, rd = torch.from_numpy(g['rays_o']), torch.from_numpy(g['rays_d'])
    N = ro.shape[0]
    pn = torch.from_numpy(np.random.default_rng(int(g['perturb_seed'])).random((N, 64), dtype=np.float32))
    un = torch.from_numpy(np.random.default_rng(int(g['pdf_seed'])).random((N, 64), dtype=np.float32))
    draws = [pn, un]
    orig_rand = torch.rand

    def fake_rand(*shape, **kw):          # renderer_df.py:363 (perturb) then :31 (sample_pdf) -- the only torch.rand calls of run()
        t = draws.pop(0)
        want = tuple(shape[0]) if len(shape) == 1 and not isinstance(shape[0], int) else tuple(shape)
        assert tuple(t.shape) == want, (t.shape, want)
        return t
    torch.rand = fake_rand
    try:
        r = net.render(ro[None], rd[None], staged=False, perturb=True, bg_color=0, ambient_ratio=1.0, shading='albedo', force_all_rays=True, **vars(opt))
    finally:
        torch.rand = orig_rand
    assert not draws
    image, ws, depth = r['image'].reshape(N, 3), r['weights_sum'].reshape(N), r['depth'].reshape(N)
    tgt = torch.from_numpy(np.random.default_rng(int(g['target_seed'])).random((N, 3), dtype=np.float32))
    loss = ((image - tgt) ** 2).mean() + 0.1 * ws.mean()
    loss.backward()
    grads = {k: v.grad.numpy() for k, v in net.named_parameters()}
    gemb = grads['encoder.embeddings']
    rows = g['gemb_rows']
    # the oracle's golden (single evaluation per point) against the reference (three passes)
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))
    print(f"[run ref] image rel {rel(g['image'], image.detach().numpy()):.2e}  ws rel {rel(g['weights_sum'], ws.detach().numpy()):.2e}  "
          f"depth rel {rel(g['depth'], depth.detach().numpy()):.2e}  loss {float(loss):.8f} vs {float(g['loss']):.8f}")
    for k in no.PARAM_KEYS[1:]:
        print(f"    grad {k:24s} rel {rel(g['g_' + k], grads[k]):.2e}")
    print(f"    grad embeddings (sampled rows) rel {rel(g['gemb_vals'], gemb[rows]):.2e}; abs-sum {np.abs(gemb).sum():.6f} vs {float(g['gemb_abs_sum']):.6f}")
    # sample_pdf on its own (renderer_df.py:15-49), stochastic and deterministic
    rng = np.random.default_rng(31)
    bins = torch.from_numpy(np.sort(rng.random((256, 63), dtype=np.float32) * 4 + 1, axis=1))
    w = torch.from_numpy(rng.random((256, 62), dtype=np.float32) ** 4)
    w[:32] = 0                                                          # pdf on its 1e-5 floor
    u = torch.from_numpy(rng.random((256, 64), dtype=np.float32))
    draws = [u]
    torch.rand = fake_rand
    try:
        sp = rdf.sample_pdf(bins, w, 64, det=False)
    finally:
        torch.rand = orig_rand
    sp_det = rdf.sample_pdf(bins, w, 64, det=True)
    np.savez_compressed(os.path.join(GOLD, 'ngp_run_ref.npz'), image=image.detach().numpy(), weights_sum=ws.detach().numpy(), depth=depth.detach().numpy(),
                        loss=float(loss.detach()), gemb_rows=rows, gemb_vals=gemb[rows], gemb_abs_sum=float(np.abs(gemb).sum()),
                        pdf_seed=31, pdf_samples=sp.numpy(), pdf_samples_det=sp_det.numpy(),
                        **{'g_' + k: grads[k] for k in no.PARAM_KEYS[1:]})


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ['unet', 'plms', 'ngp', 'vae', 'run_ref', 'eft']
    torch.set_num_threads(os.cpu_count())
    if 'unet' in which:
        gen_unet()
    if 'unet128' in which:
        gen_unet(only=('small128',))
    if 'plms' in which:
        gen_plms()
    if 'ngp' in which:
        gen_ngp()
    if 'vae' in which:
        gen_vae()
    if 'run_ref' in which:
        gen_run_ref()
    if 'eft' in which:
        gen_eft()
