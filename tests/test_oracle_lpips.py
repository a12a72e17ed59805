"""The LPIPS restatement (oracle/lpips_oracle.py) pinned on the CPU by properties of the published algorithm: identity, symmetry,
non-negativity, invariance of the per-tap normalisation to a positive rescaling of a feature map, and autograd == finite differences."""
import torch

from oracle import lpips_oracle as lo


def test_metric_properties_and_gradient():
    p = lo.make_params(seed=1)
    g = torch.Generator().manual_seed(0)
    x0 = torch.rand(1, 3, 32, 32, generator=g) * 2 - 1
    x1 = torch.rand(1, 3, 32, 32, generator=g) * 2 - 1
    d01, d10, d00 = lo.lpips(p, x0, x1), lo.lpips(p, x1, x0), lo.lpips(p, x0, x0)
    assert d01.shape == (1, 1, 1, 1) and d01.item() > 0
    assert abs(d01.item() - d10.item()) < 1e-7 and d00.item() == 0.0
    # 13 convolutions, 4 pools: feature shapes of the five taps
    f = lo.features(p, x0)
    assert [tuple(t.shape[1:]) for t in f] == [(64, 32, 32), (128, 16, 16), (256, 8, 8), (512, 4, 4), (512, 2, 2)]
    # the reference wrapper maps [0,1] -> [-1,1] (external_utils.py:37-39)
    wrap = lo.PerceptualLoss(p)
    assert torch.allclose(wrap((x0 + 1) / 2, (x1 + 1) / 2), d01, rtol=1e-5)
    # directional derivative by central differences in fp64
    pd = {k: v.double() for k, v in p.items()}
    a = x0.double().requires_grad_(True)
    lo.lpips(pd, a, x1.double()).sum().backward()
    v = torch.randn(a.shape, generator=g, dtype=torch.float64)
    eps = 1e-6
    fd = (lo.lpips(pd, x0.double() + eps * v, x1.double()) - lo.lpips(pd, x0.double() - eps * v, x1.double())).item() / (2 * eps)
    assert abs(fd - (a.grad * v).sum().item()) < 1e-6 * max(1.0, abs(fd))
