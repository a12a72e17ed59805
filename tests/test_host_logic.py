"""Host-side logic that needs no GPU: the split-K output arena, the (hi, lo) weight split, the bench's stratified max_thres sequence,
the PLMS host scalars against the oracle's schedule."""
import math

import torch


def test_zero_arena_and_meter():
    from sparsefusion_b200 import ops
    m = ops.ArenaMeter()
    assert m.take((1, 4, 4, 10)) is None and m.take((3,)) is None
    assert m.floats == 192 + 64                                  # 160 -> 192 (256-byte granules), 3 -> 64
    # the arena hands out aligned, disjoint, zero-filled views and refuses what does not fit
    buf = torch.zeros(m.floats)
    arena = ops.ZeroArena.__new__(ops.ZeroArena)
    arena.buf, arena.off = buf, 0
    a = arena.take((1, 4, 4, 10))
    b = arena.take((3,))
    assert a.shape == (1, 4, 4, 10) and b.shape == (3,) and a.data_ptr() % 256 == buf.data_ptr() % 256
    assert b.data_ptr() - a.data_ptr() == 192 * 4 and arena.take((1,)) is None
    a.fill_(1.0)
    assert float(b.sum()) == 0.0
    with ops.use_arena(arena):
        assert ops._arena is arena
    assert ops._arena is None


def test_split_packed_weight_is_exact():
    from sparsefusion_b200 import ops
    w = torch.randn(8, 64) * torch.logspace(-6, 6, 64)
    sp = ops.split_packed_weight(w)
    assert sp.shape == (2, 8, 64)
    assert torch.equal(sp[0] + sp[1], w)                                     # lo = w - hi is exact in fp32
    assert ((sp[0].view(torch.int32) & 0x1FFF) == 0).all()                  # hi carries 10 explicit mantissa bits
    assert (sp[1].abs() <= w.abs() * 2.0 ** -10 + 1e-45).all()


def test_bench_max_thres_sequence_is_stratified():
    import bench
    for k in (1, 6, 20):
        seq = bench.max_thres_sequence(3, k)
        assert len(seq) == 3 + k
        timed = sorted(seq[3:])
        assert all(abs(t - 0.99 * (i + 0.5) / k) < 1e-12 for i, t in enumerate(timed))
        assert seq == bench.max_thres_sequence(3, k)                        # seeded: identical for every leg / arm / rank
    calls = [min(int(t * 100), 50) + 1 if t >= 0.01 else 0 for t in bench.max_thres_sequence(3, 20)[3:]]
    assert abs(sum(calls) / 20 - 38.25) < 1e-9                              # the mean PLMS length the bench lines quote


def test_plms_host_scalars_match_the_schedule():
    from oracle import unet_oracle as uo
    from sparsefusion_b200 import plms
    # the conditioning log-SNR is the reference's fp32 expression (at t -> 1 fp64 would give -74.7 where the reference computes -33.9)
    assert abs(plms._log_snr(1.0) - (-33.8913)) < 1e-3
    for t, tn in [(1.0, 0.98), (0.99, 0.97), (0.5, 0.48), (0.02, 0.0), (0.3, 0.3)]:
        ls32 = uo.alpha_cosine_log_snr(torch.tensor([t, tn], dtype=torch.float32))
        assert plms._log_snr(t) == ls32[0].item() and plms._log_snr(tn) == ls32[1].item()
        ls = ls32.double()                                                  # derived scalars: fp64 arithmetic on the fp32 log-SNR values
        alpha, sigma, alpha_next, c, noise_scale = plms._step_scalars(t, tn)
        a_ref, s_ref = torch.sigmoid(ls[0]).sqrt().item(), torch.sigmoid(-ls[0]).sqrt().item()
        assert abs(alpha - a_ref) < 1e-12 and abs(sigma - s_ref) < 1e-12
        assert abs(alpha_next - torch.sigmoid(ls[1]).sqrt().item()) < 1e-12
        c_ref = -math.expm1(ls[0].item() - ls[1].item())
        assert abs(c - c_ref) < 1e-12
        if tn == 0:
            assert noise_scale == 0.0                                        # no noise is injected on the last step (plms.py:208-212)
