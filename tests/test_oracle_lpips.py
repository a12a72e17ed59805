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


def test_backbone_taps_match_torchvision_vgg16():
    """The restated feature extractor against a real third-party implementation that IS in this image: torchvision's `vgg16().features`
    (torchvision 0.26; random weights -- the pretrained ones cannot be fetched) sliced where lpips.pretrained_networks.vgg16 slices it
    ([0:4], [4:9], [9:16], [16:23], [23:30] = relu1_2, relu2_2, relu3_3, relu4_3, relu5_3).  Pins layer order, padding, pooling and the tap
    positions of the backbone half of the oracle; the `lin` heads and the normalisation stay a restatement of the published formula."""
    tv = __import__('pytest').importorskip('torchvision')
    p = lo.make_params(seed=3)
    net = tv.models.vgg16(weights=None).features.eval()
    convs = [m for m in net if isinstance(m, torch.nn.Conv2d)]
    assert len(convs) == 13
    with torch.no_grad():
        for i, m in enumerate(convs):
            m.weight.copy_(p[f'conv{i}.weight'])
            m.bias.copy_(p[f'conv{i}.bias'])
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 48, 32, generator=g) * 2 - 1
    shift = torch.tensor(lo.SHIFT).view(1, 3, 1, 1)
    scale = torch.tensor(lo.SCALE).view(1, 3, 1, 1)
    h = (x - shift) / scale
    taps = []
    with torch.no_grad():
        for lo_i, hi_i in ((0, 4), (4, 9), (9, 16), (16, 23), (23, 30)):
            for j in range(lo_i, hi_i):
                h = net[j](h)
            taps.append(h)
    ours = lo.features(p, x)
    for a, b in zip(ours, taps):
        assert a.shape == b.shape
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
