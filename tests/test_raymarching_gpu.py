"""`_raymarching` operators (C ABI) against the oracle restatement and golden vectors.

Bar: near/far bit-exact (IEEE sub/mul/div only); morton / packbits / per-ray step counts bit-exact;
composited values within 2e-4 relative (__expf vs expf).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rays(n_side=32, view=1):
    from oracle import ngp_oracle as no
    return no.camera_rays(no.circle_cameras(8)[view], n_side, n_side)


def test_near_far_bit_exact():
    from oracle import ngp_oracle as no
    from sparsefusion_b200 import raymarching as rm
    ro, rd = _rays(64)
    rng = np.random.default_rng(0)
    ro2 = np.concatenate([ro, rng.standard_normal((4096, 3)).astype(np.float32) * 6])
    rd2 = np.concatenate([rd, rng.standard_normal((4096, 3)).astype(np.float32)])
    aabb = np.array([-4, -4, -4, 4, 4, 4], np.float32)
    for min_near in (0.1, 0.2):
        n, f = rm.near_far_from_aabb(torch.from_numpy(ro2).cuda(), torch.from_numpy(rd2).cuda(), torch.from_numpy(aabb).cuda(), min_near)
        no_, fo = no.near_far_from_aabb(ro2, rd2, aabb, min_near)
        assert np.array_equal(n.cpu().numpy(), no_) and np.array_equal(f.cpu().numpy(), fo)
    assert (no_ == np.finfo(np.float32).max).any(), 'test should contain rays that miss the box'


def test_morton_roundtrip_and_packbits():
    from oracle import ngp_oracle as no
    from sparsefusion_b200 import raymarching as rm
    rng = np.random.default_rng(1)
    coords = rng.integers(0, 128, size=(100000, 3)).astype(np.int32)
    idx = rm.morton3D(torch.from_numpy(coords).cuda())
    assert np.array_equal(idx.cpu().numpy(), no.morton3D(coords))
    back = rm.morton3D_invert(idx)
    assert np.array_equal(back.cpu().numpy(), coords)
    ar = np.arange(128 ** 3, dtype=np.int32)  # full 128^3 domain: a bijection
    inv = rm.morton3D_invert(torch.from_numpy(ar).cuda())
    assert np.array_equal(rm.morton3D(inv).cpu().numpy(), ar)
    grid = rng.standard_normal((3, 128 ** 3)).astype(np.float32)
    bits = rm.packbits(torch.from_numpy(grid).cuda(), 0.3)
    assert np.array_equal(bits.cpu().numpy(), no.packbits(grid, 0.3))
    assert rm.morton3D(torch.zeros(0, 3, dtype=torch.int32, device='cuda')).shape == (0,)


def test_sph_from_ray():
    from oracle import ngp_oracle as no
    from sparsefusion_b200 import raymarching as rm
    ro, rd = _rays(32)
    c = rm.sph_from_ray(torch.from_numpy(ro * 0.1).cuda(), torch.from_numpy(rd).cuda(), 8.0)
    np.testing.assert_allclose(c.cpu().numpy(), no.sph_from_ray(ro * 0.1, rd, 8.0), rtol=1e-4, atol=1e-5)


def test_march_train_and_composite_vs_oracle_and_golden(golden_dir):
    from oracle import ngp_oracle as no
    from sparsefusion_b200 import raymarching as rm, _raymarching as be
    g = np.load(f'{golden_dir}/ngp_march.npz')
    ro, rd, bitfield, noises = g['rays_o'], g['rays_d'], g['bitfield'], g['noises']
    dev = lambda a: torch.from_numpy(a).cuda()
    nears, fars = rm.near_far_from_aabb(dev(ro), dev(rd), dev(np.array([-4, -4, -4, 4, 4, 4], np.float32)), 0.2)
    assert np.array_equal(nears.cpu().numpy(), g['nears'])
    N, max_steps = ro.shape[0], 256
    M = N * max_steps
    xyzs, dirs, deltas = (torch.zeros(M, 3, device='cuda'), torch.zeros(M, 3, device='cuda'), torch.zeros(M, 2, device='cuda'))
    rays = torch.empty(N, 3, dtype=torch.int32, device='cuda')
    counter = torch.zeros(2, dtype=torch.int32, device='cuda')
    be.march_rays_train(dev(ro), dev(rd), dev(bitfield), 4.0, 0.0, max_steps, N, 3, 128, M, nears, fars, xyzs, dirs, deltas, rays, counter, dev(noises))
    rays_h = rays.cpu().numpy()
    order = np.argsort(rays_h[:, 0])
    assert np.array_equal(rays_h[order, 0], np.arange(N))
    # per-ray step counts are the integer contract (point offsets depend on atomic arrival order)
    assert np.array_equal(rays_h[order, 2], g['rays'][:, 2]), 'per-ray sample counts differ from the golden vector'
    assert int(counter[0]) == int(g['n_points']) and int(counter[1]) == N
    xo, do, lo, ro_tab, _ = no.march_rays_train(ro, rd, 4.0, bitfield, 3, 128, g['nears'], g['fars'], noises, 0.0, max_steps)
    xh, lh = xyzs.cpu().numpy(), deltas.cpu().numpy()
    for n in np.random.default_rng(3).choice(N, 64, replace=False):  # per-ray comparison through the rays table
        _, off_d, cnt = rays_h[order][n]
        _, off_o, cnt_o = ro_tab[n]
        assert cnt == cnt_o
        np.testing.assert_array_equal(xh[off_d:off_d + cnt], xo[off_o:off_o + cnt])
        np.testing.assert_array_equal(lh[off_d:off_d + cnt], lo[off_o:off_o + cnt])
    # composite forward / backward on the device-marched points
    m = int(counter[0])
    rng = np.random.default_rng(4)
    sig = (rng.random(m, dtype=np.float32) * 20).astype(np.float32)
    rgb = rng.random((m, 3), dtype=np.float32)
    s_t, c_t = dev(sig).requires_grad_(True), dev(rgb).requires_grad_(True)
    ws, depth, image = rm.composite_rays_train(s_t, c_t, deltas[:m].contiguous(), rays, 1e-4)
    wso, deptho, imo = no.composite_rays_train_forward(sig, rgb, lh[:m], rays_h, 1e-4)
    np.testing.assert_allclose(ws.detach().cpu().numpy(), wso, rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(image.detach().cpu().numpy(), imo, rtol=2e-4, atol=1e-6)
    np.testing.assert_allclose(depth.detach().cpu().numpy(), deptho, rtol=2e-4, atol=1e-5)
    gws, gim = rng.standard_normal(N).astype(np.float32), rng.standard_normal((N, 3)).astype(np.float32)
    (ws * dev(gws)).sum().add((image * dev(gim)).sum()).backward()
    gso, gco = no.composite_rays_train_backward(gws, gim, sig, rgb, lh[:m], rays_h, wso, imo, 1e-4)
    np.testing.assert_allclose(c_t.grad.cpu().numpy(), gco, rtol=3e-4, atol=1e-6)
    np.testing.assert_allclose(s_t.grad.cpu().numpy(), gso, rtol=2e-3, atol=2e-4)


def test_march_and_composite_inference_loop():
    """run the reference's eval loop (renderer_df.py:521-557) with device operators and with the oracle; same alive sets"""
    from oracle import ngp_oracle as no
    from sparsefusion_b200 import raymarching as rm
    ro, rd = _rays(32, view=3)
    N = ro.shape[0]
    rng = np.random.default_rng(6)
    grid = (rng.random((3, 128 ** 3), dtype=np.float32) < 0.02).astype(np.float32)
    bitfield = no.packbits(grid, 0.5)
    aabb = np.array([-4, -4, -4, 4, 4, 4], np.float32)
    nears, fars = no.near_far_from_aabb(ro, rd, aabb, 0.2)
    dev = lambda a: torch.from_numpy(a).cuda()
    sigma_fn = lambda x: (np.abs(np.sin(x.sum(-1) * 3)) * 8).astype(np.float32)
    rgb_fn = lambda x: (0.5 + 0.5 * np.cos(x * 2)).astype(np.float32)
    # oracle
    ws_o, d_o, im_o = np.zeros(N, np.float32), np.zeros(N, np.float32), np.zeros((N, 3), np.float32)
    alive_o, t_o = np.arange(N, dtype=np.int32), nears.copy()
    # device
    ws_d, d_d, im_d = torch.zeros(N, device='cuda'), torch.zeros(N, device='cuda'), torch.zeros(N, 3, device='cuda')
    alive_d, t_d = torch.arange(N, dtype=torch.int32, device='cuda'), dev(nears.copy())
    step = 0
    while step < 256:
        n_alive = alive_o.shape[0]
        assert alive_d.shape[0] == n_alive
        if n_alive == 0:
            break
        n_step = max(min(N // n_alive, 8), 1)
        xo, _, lo = no.march_rays(n_alive, n_step, alive_o, t_o, ro, rd, 4.0, bitfield, 3, 128, nears, fars, np.zeros(n_alive, np.float32), 0.0, 256, 128)
        xd, _, ld = rm.march_rays(n_alive, n_step, alive_d, t_d, dev(ro), dev(rd), 4.0, dev(bitfield), 3, 128, dev(nears), dev(fars), 128, False, 0, 256)
        np.testing.assert_array_equal(xd.cpu().numpy(), xo)
        np.testing.assert_array_equal(ld.cpu().numpy(), lo)
        so, co = sigma_fn(xo), rgb_fn(xo)
        no.composite_rays(n_alive, n_step, alive_o, t_o, so, co, lo, ws_o, d_o, im_o, 1e-2)
        rm.composite_rays(n_alive, n_step, alive_d, t_d, dev(so), dev(co), ld, ws_d, d_d, im_d, 1e-2)
        assert np.array_equal(alive_d.cpu().numpy(), alive_o)
        alive_o = np.ascontiguousarray(alive_o[alive_o >= 0])
        alive_d = alive_d[alive_d >= 0]
        step += n_step
    np.testing.assert_allclose(ws_d.cpu().numpy(), ws_o, rtol=3e-4, atol=1e-6)
    np.testing.assert_allclose(im_d.cpu().numpy(), im_o, rtol=3e-4, atol=1e-6)
