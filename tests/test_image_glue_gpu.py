"""Ray generation and the fused image-space losses (SURVEY §8f row 2, C ABI section 6) against the harness' ray function and
against torch autograd through the expressions of sparsefusion/distillation.py:210-234, :287-344 (fp64)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _huber(x, y, s=0.1):   # utils/common_utils.py:183-190
    return ((1 + (x - y) ** 2 / s ** 2).clamp(1e-4).sqrt() - 1) * s


def _rel(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm().clamp(min=1e-30)).item()


@pytest.mark.parametrize('hw', [(128, 128), (37, 53), (256, 256)])
def test_rays_match_the_harness_camera_function(hw):
    from oracle import ngp_oracle as no
    from sparsefusion_b200 import image_glue as glue
    h, w = hw
    for cam in no.circle_cameras(5)[:3]:
        o_ref, d_ref = no.camera_rays(cam, h, w, 4.0)
        o, d = glue.rays_from_camera(torch.from_numpy(cam[0]).cuda(), torch.from_numpy(cam[1]).cuda(), h, w, 4.0)
        assert np.array_equal(o.cpu().numpy(), o_ref)
        assert np.abs(d.cpu().numpy() - d_ref).max() < 2e-7          # same expression; numpy's matmul may order / fuse the 3-term sum differently


@pytest.mark.parametrize('hw', [64, 128])
def test_photometric_loss_and_gradients(hw):
    from sparsefusion_b200 import image_glue as glue
    g = torch.Generator(device='cuda').manual_seed(1)
    n = hw * hw
    img = torch.rand(n, 3, device='cuda', generator=g)
    sil = torch.rand(n, device='cuda', generator=g)
    rgb = torch.rand(3, 2 * hw, 2 * hw, device='cuda', generator=g)
    mask = (torch.rand(1, 2 * hw, 2 * hw, device='cuda', generator=g) > 0.5).float()
    lc, ls, lo = 1.0, 0.7, 1e-3
    loss, g_img, g_sil = glue.photometric_loss(img, sil, rgb, mask, hw, hw, 2, lc, ls, lo)
    a = img.double().requires_grad_(True)
    b = sil.double().requires_grad_(True)
    image = a.reshape(1, hw, hw, 3).permute(0, 3, 1, 2)
    s = b.reshape(1, 1, hw, hw)
    ref = (lc * _huber(image, F.interpolate(rgb[None].double(), scale_factor=0.5)).abs().mean()
           + ls * _huber(s, F.interpolate(mask[None].double(), scale_factor=0.5)).abs().mean() + lo * torch.sqrt(s ** 2 + .01).mean())
    ref.backward()
    assert abs(loss.item() - ref.item()) < 2e-6 * abs(ref.item())
    assert _rel(g_img, a.grad) < 2e-6 and _rel(g_sil, b.grad) < 2e-6


@pytest.mark.parametrize('mode', ['sds', 'eft'])
@pytest.mark.parametrize('hw', [16, 128])
def test_fusion_loss_upsample_and_adjoint(mode, hw):
    from sparsefusion_b200 import image_glue as glue
    g = torch.Generator(device='cuda').manual_seed(2)
    n = hw * hw
    img = torch.rand(n, 3, device='cuda', generator=g)
    sil = torch.rand(n, device='cuda', generator=g)
    target = torch.rand(3, 2 * hw, 2 * hw, device='cuda', generator=g) * 0.3
    lc, ls, lo, wgt = 1.0, 0.7, 1e-3, 0.37
    up = glue.upsample2x_render(img, sil, hw, hw)
    a = img.double().requires_grad_(True)
    b = sil.double().requires_grad_(True)
    image = F.interpolate(a.reshape(1, hw, hw, 3).permute(0, 3, 1, 2), scale_factor=2, mode='bilinear')
    s = F.interpolate(b.reshape(1, 1, hw, hw), scale_factor=2, mode='bilinear')
    assert _rel(up[:3], image[0]) < 1e-6 and _rel(up[3], s[0, 0]) < 1e-6
    t = target[None].double()
    if mode == 'sds':
        fl = wgt * (image - t).abs().mean()
    else:
        fl = lc * _huber(image, t).abs().mean() + ls * _huber(s, (t.mean(dim=1, keepdim=True) > .1).double()).abs().mean()
    ref = fl + lo * torch.sqrt(s ** 2 + .01).mean()
    ref.backward()
    loss, g_img, g_sil = glue.fusion_loss(up, target, hw, hw, mode, wgt, lc, ls, lo)
    assert abs(loss.item() - ref.item()) < 3e-6 * abs(ref.item())
    assert _rel(g_img, a.grad) < 3e-6 and _rel(g_sil, b.grad) < 3e-6
