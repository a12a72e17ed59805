"""Grid-encoder operators (C ABI, through the `_gridencoder` drop-in) against the oracle and the golden vectors.

Bar: embedding-row indices bit-exact; interpolated values within 1e-5 relative (fp32, different fma
contraction at most); backward scatter within 1e-4 (atomic order).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _device_scales(geo):
    from sparsefusion_b200 import _lib as lib
    s = torch.empty(geo['L'], device='cuda')
    lib.call('sfb_grid_level_scales', geo['L'], float(geo['S']), geo['H'], lib.fptr(s), lib.stream())
    return s.cpu().numpy()


def _device_rows(x, geo, offsets=None, L=None, S=None, H=None, gridtype=None):
    from sparsefusion_b200 import _lib as lib
    offsets = geo['offsets'] if offsets is None else offsets
    L = geo['L'] if L is None else L
    xt = torch.from_numpy(x).cuda()
    off = torch.from_numpy(offsets).cuda()
    rows = torch.empty(L, x.shape[0], 8, dtype=torch.int32, device='cuda')
    lib.call('sfb_grid_corner_rows', lib.fptr(xt), lib.iptr(off), lib.iptr(rows), x.shape[0], 3, L, float(geo['S'] if S is None else S),
             geo['H'] if H is None else H, geo['gridtype'] if gridtype is None else gridtype, 0, lib.stream())
    return rows.cpu().numpy()


def test_level_scales_device_vs_host():
    """informative: device exp2f vs host libm.  The index contract is defined by the DEVICE value
    (gridencoder.cu:125); the oracle is fed the device scales wherever they differ."""
    from oracle import ngp_oracle as no
    geo = no.live_geometry()
    dev, host = _device_scales(geo), no.level_scales_host(geo['L'], geo['S'], geo['H'])
    diff = np.abs(dev.view(np.int32) - host.view(np.int32))
    print('level scale ulp differences device-vs-host:', diff.tolist())
    assert diff.max() <= 4
    assert dev[0] == 15.0 and dev[5] == 127.0 and dev[15] == 8191.0 or diff.max() > 0


def test_corner_rows_bit_exact_golden_and_random(golden_dir):
    from oracle import ngp_oracle as no
    geo = no.live_geometry()
    g = np.load(f'{golden_dir}/ngp_grid.npz')
    scales = _device_scales(geo)
    x = g['x']
    _, _, rows_o = no.grid_encode_forward(x, np.zeros((int(geo['offsets'][-1]), 2), np.float32), geo['offsets'], geo['S'], geo['H'], 1, False,
                                          False, scales, True)
    rows_d = _device_rows(x, geo)
    assert np.array_equal(rows_d, rows_o), f'{(rows_d != rows_o).sum()} corner rows differ'
    if np.array_equal(scales, g['level_scales']):
        assert np.array_equal(rows_d, g['rows'])
    # a million random points
    xr = np.random.default_rng(0).random((1 << 20, 3), dtype=np.float32)
    _, _, ro = no.grid_encode_forward(xr, np.zeros((int(geo['offsets'][-1]), 2), np.float32), geo['offsets'], geo['S'], geo['H'], 1, False, False,
                                      scales, True)
    assert np.array_equal(_device_rows(xr, geo), ro)


def test_hash_grid_rows_bit_exact(golden_dir):
    from oracle import ngp_oracle as no
    g = np.load(f'{golden_dir}/ngp_grid_hash.npz')
    offs = g['offsets']
    S = float(np.log2(2.0))
    geo = dict(S=S, H=16, gridtype=0, L=8, offsets=offs)
    scales = _device_scales(geo)
    rows_d = _device_rows(g['x'], geo)
    _, _, rows_o = no.grid_encode_forward(g['x'], np.zeros((int(offs[-1]), 2), np.float32), offs, S, 16, 0, False, False, scales, True)
    assert np.array_equal(rows_d, rows_o)
    assert np.array_equal(scales, g['level_scales']) and np.array_equal(rows_d, g['rows'])


def test_forward_backward_values_vs_oracle(golden_dir):
    from oracle import ngp_oracle as no
    from sparsefusion_b200 import _gridencoder as be
    geo = no.live_geometry()
    scales = _device_scales(geo)
    p = no.make_field_params(seed=0)
    emb = p['encoder.embeddings']
    g = np.load(f'{golden_dir}/ngp_grid.npz')
    x = g['x']
    B, L, C, D = x.shape[0], geo['L'], geo['C'], 3
    xt, et, off = torch.from_numpy(x).cuda(), emb.cuda(), torch.from_numpy(geo['offsets']).cuda()
    out = torch.empty(L, B, C, device='cuda')
    dy = torch.empty(B, L * D * C, device='cuda')
    be.grid_encode_forward(xt, et, off, out, B, D, C, L, geo['S'], geo['H'], dy, 1, False)
    out_o, dy_o, _ = no.grid_encode_forward(x, emb.numpy(), geo['offsets'], geo['S'], geo['H'], 1, False, True, scales)
    np.testing.assert_allclose(out.cpu().numpy(), out_o, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(dy.cpu().numpy(), dy_o, rtol=1e-4, atol=1e-3)
    if np.array_equal(scales, g['level_scales']):
        np.testing.assert_allclose(out.cpu().numpy(), g['out'], rtol=1e-5, atol=1e-6)
    grad = np.random.default_rng(5).standard_normal(out_o.shape, dtype=np.float32)
    ge = torch.zeros_like(et)
    gi = torch.zeros(B, D, device='cuda')
    be.grid_encode_backward(torch.from_numpy(grad).cuda(), xt, et, off, ge, B, D, C, L, geo['S'], geo['H'], dy, gi, 1, False)
    ge_o, gi_o = no.grid_encode_backward(grad, x, geo['offsets'], emb.shape[0], geo['S'], geo['H'], 1, False, dy_o, scales)
    np.testing.assert_allclose(ge.cpu().numpy(), ge_o, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gi.cpu().numpy(), gi_o, rtol=1e-3, atol=2e-2)
    if np.array_equal(scales, g['level_scales']):
        np.testing.assert_allclose(ge.cpu().numpy()[g['ge_rows']], g['ge_vals'], rtol=1e-4, atol=1e-5)


def test_edge_cases_empty_and_out_of_range():
    from sparsefusion_b200.gridencoder import GridEncoder
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16, desired_resolution=8192,
                      gridtype='tiled').cuda()
    enc.embeddings.data.uniform_(-0.5, 0.5)
    assert enc(torch.zeros(0, 3, device='cuda'), bound=4).shape == (0, 32)
    out = enc(torch.tensor([[4.5, 0.0, 0.0], [0.0, -4.0001, 0.0]], device='cuda'), bound=4)
    assert (out == 0).all(), 'points outside [0,1]^3 encode to zeros (gridencoder.cu:98-122)'
    with pytest.raises(RuntimeError):
        from sparsefusion_b200 import _gridencoder as be
        be.grid_encode_forward(torch.zeros(4, 3), enc.embeddings.data, enc.offsets, torch.zeros(16, 4, 2), 4, 3, 2, 16, 0.6, 16, None, 1, False)


def test_module_autograd_partition_of_unity_and_adjoint():
    """properties at full size: with all embeddings == 1 every level interpolates to exactly ~1 (weights sum to 1);
    <grad, enc(e)> == <backward(grad), e> (the backward is the adjoint of the forward, which is linear in the table)."""
    from sparsefusion_b200.gridencoder import GridEncoder
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16, desired_resolution=8192,
                      gridtype='tiled').cuda()
    x = (torch.rand(1 << 18, 3, device='cuda') * 2 - 1) * 4
    enc.embeddings.data.fill_(1.0)
    y = enc(x, bound=4)
    assert (y - 1).abs().max().item() < 1e-5
    enc.embeddings.data.uniform_(-0.5, 0.5)
    y = enc(x, bound=4)
    gr = torch.randn_like(y)
    (y * gr).sum().backward()
    lhs = (y.detach().double() * gr.double()).sum().item()
    rhs = (enc.embeddings.grad.double() * enc.embeddings.data.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-4 * abs(lhs) + 1e-3
