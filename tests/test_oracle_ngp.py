"""The C restatement of the NGP operators (oracle/ngp_oracle.c) against its committed golden vectors and
against size-independent properties.  (Cross-validation against the reference's CUDA sources happens on the GPU
box: tests/test_ref_cuda_gpu.py.)"""
import numpy as np
import pytest
import torch


def test_geometry_matches_survey_appendix_b():
    from oracle import ngp_oracle as no
    geo = no.live_geometry()
    off = geo['offsets']
    assert off[-1] == 929_336 and list(np.diff(off)[:4]) == [4920, 17576, 54872, 65536] and abs(geo['S'] - 0.6) < 1e-6
    sc = no.level_scales_host(16, geo['S'], 16)
    assert sc[0] == 15.0 and sc[5] == 127.0 and abs(sc[15] - 8191.0) < 1e-2


def test_grid_forward_backward_golden(golden_dir):
    from oracle import ngp_oracle as no
    geo = no.live_geometry()
    g = np.load(f'{golden_dir}/ngp_grid.npz')
    emb = no.make_field_params(seed=0)['encoder.embeddings'].numpy()
    out, dy, rows = no.grid_encode_forward(g['x'], emb, geo['offsets'], geo['S'], geo['H'], 1, False, True, g['level_scales'], True)
    assert np.array_equal(rows, g['rows']) and np.array_equal(out, g['out'])
    grad = np.random.default_rng(int(g['grad_seed'])).standard_normal(out.shape, dtype=np.float32)
    ge, gi = no.grid_encode_backward(grad, g['x'], geo['offsets'], emb.shape[0], geo['S'], geo['H'], 1, False, dy, g['level_scales'])
    assert np.array_equal(ge[g['ge_rows']], g['ge_vals']) and abs(np.abs(ge).sum() - float(g['ge_abs_sum'])) < 1e-3 * float(g['ge_abs_sum'])
    assert np.allclose(gi, g['gi'])


def test_tiled_index_quirk_drops_z_from_level_7():
    """gridencoder.cu:60: the stride loop stops once stride > rows, so for (res+1)^2 > 65536 z never enters the index"""
    from oracle import ngp_oracle as no
    geo = no.live_geometry()
    x = np.array([[0.3, 0.6, 0.1], [0.3, 0.6, 0.9]], np.float32)
    _, _, rows = no.grid_encode_forward(x, np.zeros((int(geo['offsets'][-1]), 2), np.float32), geo['offsets'], geo['S'], geo['H'], 1, False, False, None, True)
    assert (rows[7:, 0, :4] == rows[7:, 1, :4]).all() and (rows[:6, 0] != rows[:6, 1]).any()


def test_grid_properties_partition_of_unity_linearity_adjoint():
    from oracle import ngp_oracle as no
    geo = no.live_geometry()
    rng = np.random.default_rng(0)
    x = rng.random((2000, 3), dtype=np.float32)
    rows = int(geo['offsets'][-1])
    ones, _, _ = no.grid_encode_forward(x, np.ones((rows, 2), np.float32), geo['offsets'], geo['S'], geo['H'])
    assert np.abs(ones - 1).max() < 1e-5
    a, b = rng.standard_normal((rows, 2)).astype(np.float32), rng.standard_normal((rows, 2)).astype(np.float32)
    fa, _, _ = no.grid_encode_forward(x, a, geo['offsets'], geo['S'], geo['H'])
    fb, _, _ = no.grid_encode_forward(x, b, geo['offsets'], geo['S'], geo['H'])
    fab, _, _ = no.grid_encode_forward(x, a + 2 * b, geo['offsets'], geo['S'], geo['H'])
    assert np.abs(fab - (fa + 2 * fb)).max() < 1e-4
    gr = rng.standard_normal(fa.shape).astype(np.float32)
    ge, _ = no.grid_encode_backward(gr, x, geo['offsets'], rows, geo['S'], geo['H'])
    assert abs((fa.astype(np.float64) * gr).sum() - (ge.astype(np.float64) * a).sum()) < 1e-3 * abs((fa.astype(np.float64) * gr).sum()) + 1e-2


def test_march_and_run_golden(golden_dir):
    from oracle import ngp_oracle as no
    g = np.load(f'{golden_dir}/ngp_march.npz')
    nears, fars = no.near_far_from_aabb(g['rays_o'], g['rays_d'], np.array([-4, -4, -4, 4, 4, 4], np.float32), 0.2)
    assert np.array_equal(nears, g['nears']) and np.array_equal(fars, g['fars'])
    _, _, _, rays, counter = no.march_rays_train(g['rays_o'], g['rays_d'], 4.0, g['bitfield'], 3, 128, nears, fars, g['noises'], 0.0, 256)
    assert np.array_equal(rays, g['rays']) and int(counter[0]) == int(g['n_points'])
    field = no.Field(no.make_field_params(seed=0))
    res = no.run_cuda_train(field, torch.from_numpy(g['rays_o']), torch.from_numpy(g['rays_d']), g['bitfield'], noises=g['noises'])
    assert np.allclose(res['image'].detach().numpy(), g['image'], atol=1e-6)
    ev = no.run_cuda_eval(field, torch.from_numpy(g['rays_o']), torch.from_numpy(g['rays_d']), g['bitfield'])
    assert np.allclose(ev['image'].numpy(), g['eval_image'], atol=1e-6)
    # train and eval marching see the same occupancy -> same picture up to the T threshold used for early stop
    assert np.abs(ev['image'].numpy() - res['image'].detach().numpy()).max() < 5e-2


def test_run_golden_and_gradient(golden_dir):
    from oracle import ngp_oracle as no
    g = np.load(f'{golden_dir}/ngp_run.npz')
    N = g['rays_o'].shape[0]
    pn = torch.from_numpy(np.random.default_rng(int(g['perturb_seed'])).random((N, 64), dtype=np.float32))
    un = torch.from_numpy(np.random.default_rng(int(g['pdf_seed'])).random((N, 64), dtype=np.float32))
    params = {k: v.clone().requires_grad_(True) for k, v in no.make_field_params(seed=0).items()}
    r = no.run(no.Field(params), torch.from_numpy(g['rays_o']), torch.from_numpy(g['rays_d']), perturb_noise=pn, pdf_noise=un)
    assert np.allclose(r['image'].detach().numpy(), g['image'], atol=1e-6)
    z = r['z_vals'].numpy()
    assert (np.diff(z, axis=1) >= 0).all(), 'merged coarse+fine samples must be sorted'
    tgt = torch.from_numpy(np.random.default_rng(int(g['target_seed'])).random((N, 3), dtype=np.float32))
    loss = ((r['image'] - tgt) ** 2).mean() + 0.1 * r['weights_sum'].mean()
    loss.backward()
    assert abs(float(loss) - float(g['loss'])) < 1e-6
    assert np.allclose(params['encoder.embeddings'].grad.numpy()[g['gemb_rows']], g['gemb_vals'], atol=1e-7, rtol=1e-4)
    assert np.allclose(params['sigma_net.net.2.weight'].grad.numpy(), g['g_sigma_net.net.2.weight'], atol=1e-7, rtol=1e-4)


def test_morton_bijection_and_packbits_roundtrip():
    from oracle import ngp_oracle as no
    ar = np.arange(128 ** 3, dtype=np.int32)
    assert np.array_equal(no.morton3D(no.morton3D_invert(ar)), ar)
    grid = np.random.default_rng(2).standard_normal(4096).astype(np.float32)
    bits = np.unpackbits(no.packbits(grid, 0.1), bitorder='little')
    assert np.array_equal(bits.astype(bool), grid > 0.1)


def test_run_restatement_equals_the_references_own_three_pass_run(golden_dir):
    """ngp_run_ref.npz is the output of the REFERENCE's own NeRFNetwork.run / sample_pdf / MLP / trunc_exp executed on CPU (oracle/gen_golden.py
    gen_run_ref: renderer_df.py:310-468 with its three field passes :373,:398,:424).  The oracle's run() evaluates the field once at the sorted
    samples; this test pins that restructuring -- values AND every parameter gradient -- to the literal reference."""
    from oracle import ngp_oracle as no
    g, ref = np.load(f'{golden_dir}/ngp_run.npz'), np.load(f'{golden_dir}/ngp_run_ref.npz')
    N = g['rays_o'].shape[0]
    pn = torch.from_numpy(np.random.default_rng(int(g['perturb_seed'])).random((N, 64), dtype=np.float32))
    un = torch.from_numpy(np.random.default_rng(int(g['pdf_seed'])).random((N, 64), dtype=np.float32))
    params = {k: v.clone().requires_grad_(True) for k, v in no.make_field_params(seed=0).items()}
    r = no.run(no.Field(params), torch.from_numpy(g['rays_o']), torch.from_numpy(g['rays_d']), perturb_noise=pn, pdf_noise=un)
    for k in ('image', 'weights_sum', 'depth'):
        assert np.allclose(r[k].detach().numpy(), ref[k], atol=1e-6), k
    tgt = torch.from_numpy(np.random.default_rng(int(g['target_seed'])).random((N, 3), dtype=np.float32))
    loss = ((r['image'] - tgt) ** 2).mean() + 0.1 * r['weights_sum'].mean()
    loss.backward()
    assert abs(float(loss.detach()) - float(ref['loss'])) < 1e-6
    rel = lambda a, b: float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))
    for k in no.PARAM_KEYS[1:]:
        assert rel(params[k].grad.numpy(), ref['g_' + k]) < 1e-5, k
    gemb = params['encoder.embeddings'].grad.numpy()
    assert rel(gemb[ref['gemb_rows']], ref['gemb_vals']) < 1e-5
    assert abs(np.abs(gemb).sum() - float(ref['gemb_abs_sum'])) < 1e-4 * float(ref['gemb_abs_sum'])


def test_sample_pdf_equals_the_references(golden_dir):
    """renderer_df.py:15-49 executed from the reference (stochastic u injected, and det=True) against oracle.sample_pdf"""
    from oracle import ngp_oracle as no
    ref = np.load(f'{golden_dir}/ngp_run_ref.npz')
    rng = np.random.default_rng(int(ref['pdf_seed']))
    bins = torch.from_numpy(np.sort(rng.random((256, 63), dtype=np.float32) * 4 + 1, axis=1))
    w = torch.from_numpy(rng.random((256, 62), dtype=np.float32) ** 4)
    w[:32] = 0
    u = torch.from_numpy(rng.random((256, 64), dtype=np.float32))
    assert np.array_equal(no.sample_pdf(bins, w, 64, det=False, u=u).numpy(), ref['pdf_samples'])
    assert np.array_equal(no.sample_pdf(bins, w, 64, det=True).numpy(), ref['pdf_samples_det'])


def test_literal_three_pass_run_of_the_gpu_reference_leg(golden_dir):
    """oracle/gpu_reference.run_three_pass (what bench.py's eager-GPU baseline times, there with the reference's CUDA operators) restates
    renderer_df.run literally; with the C stand-ins for the two CUDA operators it reproduces the reference's own output bit for bit on CPU"""
    from oracle import gpu_reference as gr, ngp_oracle as no
    g, ref = np.load(f'{golden_dir}/ngp_run.npz'), np.load(f'{golden_dir}/ngp_run_ref.npz')
    N = g['rays_o'].shape[0]
    pn = torch.from_numpy(np.random.default_rng(int(g['perturb_seed'])).random((N, 64), dtype=np.float32))
    un = torch.from_numpy(np.random.default_rng(int(g['pdf_seed'])).random((N, 64), dtype=np.float32))
    params = {k: v.clone().requires_grad_(True) for k, v in no.make_field_params(seed=0).items()}
    f = no.Field(params)
    f.density = lambda x: dict(zip(('sigma', 'albedo'), f.common_forward(x)))

    def nf(ro, rd, aabb, mn):
        n, fa = no.near_far_from_aabb(ro.numpy(), rd.numpy(), aabb.numpy(), mn)
        return torch.from_numpy(n), torch.from_numpy(fa)
    r = gr.run_three_pass(f, torch.from_numpy(g['rays_o']), torch.from_numpy(g['rays_d']), perturb_noise=pn, pdf_noise=un, near_far_fn=nf)
    for k in ('image', 'weights_sum', 'depth'):
        assert np.allclose(r[k].detach().numpy(), ref[k], atol=1e-7), k
    tgt = torch.from_numpy(np.random.default_rng(int(g['target_seed'])).random((N, 3), dtype=np.float32))
    (((r['image'] - tgt) ** 2).mean() + 0.1 * r['weights_sum'].mean()).backward()
    for k in no.PARAM_KEYS[1:]:
        assert np.allclose(params[k].grad.numpy(), ref['g_' + k], rtol=1e-4, atol=1e-8), k
