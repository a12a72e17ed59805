"""Epipolar Feature Transformer mirror, CPU side: state_dict layout of the reference (sparsefusion/eft.py built as
utils/load_model.py:34 does), the camera / embedding restatements, and the no-CPU-fallback rule.  The arithmetic is checked on the GPU box
against tests/golden/eft.npz, minted by the reference's own class (tests/test_eft_gpu.py)."""
import numpy as np
import pytest
import torch


def test_state_dict_layout_and_golden_metadata(golden_dir):
    from sparsefusion_b200.eft import EpipolarFeatureTransformer
    m = EpipolarFeatureTransformer(use_r=True, encoder='resnet18', return_features=True, remove_unused_layers=False)
    sd = m.state_dict()
    g = np.load(f'{golden_dir}/eft.npz')
    assert len(sd) == int(g['n_tensors']) == 278          # oracle/gen_golden.py asserted key-for-key equality with the reference module when minting
    for k, shape in (('encoder_model.conv1.weight', (64, 3, 7, 7)), ('encoder_model.layer3.0.downsample.0.weight', (256, 128, 1, 1)),
                     ('encoder_model.layer4.1.bn2.running_var', (512,)), ('encoder_model.fc.weight', (1000, 512)), ('t1.pre.0.weight', (256, 606)),
                     ('t2.pre.0.weight', (256, 425)), ('t3.pre.0.weight', (256, 412)), ('t1.encoder.layers.3.self_attn.in_proj_weight', (768, 256)),
                     ('t2.encoder.layers.0.linear1.weight', (256, 256)), ('t2_attn.weight', (1, 256)), ('t3_attn.bias', (1,)), ('color_layer.0.weight', (3, 256))):
        assert tuple(sd[k].shape) == shape, k
    assert m.harmonic_embedding.get_output_dim(6) == 78 and m.harmonic_embedding.get_output_dim(1) == 13 and m.feat_size == 512
    assert EpipolarFeatureTransformer(use_r=True, encoder='resnet18', return_features=True, remove_unused_layers=True).state_dict().keys() < sd.keys()
    with pytest.raises(NotImplementedError):
        EpipolarFeatureTransformer(encoder='lite')


def test_camera_and_embedding_restatements_agree_with_the_oracles():
    from oracle import eft_oracle as eo
    from sparsefusion_b200 import eft
    cams = eo.look_at_cameras(3)
    plain = type('C', (), {})()                            # an object with the four tensors only -> the product's own restatement is used
    plain.R, plain.T, plain.focal_length, plain.principal_point = cams.R, cams.T, cams.focal_length, cams.principal_point
    pts = torch.randn(1, 50, 3)
    assert torch.allclose(eft.transform_points_ndc(plain, pts), cams.transform_points_ndc(pts), atol=1e-6)
    assert torch.allclose(eft.camera_center(plain), cams.get_camera_center(), atol=1e-6)
    c = cams.get_camera_center()                           # centres project onto the optical axis: x_view = 0 at the centre, and the origin is in front
    assert torch.allclose(torch.bmm(c[:, None, :], cams.R)[:, 0] + cams.T, torch.zeros(3, 3), atol=1e-5)
    assert (cams.transform_points_ndc(torch.zeros(1, 1, 3))[..., 2] > 0).all()
    x = torch.randn(7, 6)
    e = eft.HarmonicEmbedding(6, 1.0)(x)
    assert e.shape == (7, 78) and torch.allclose(e[:, :6], torch.sin(x[:, :1] * (2.0 ** torch.arange(6))), atol=1e-6) and torch.equal(e[:, -6:], x)
    rb = eo.RayBundle(torch.randn(4, 3), torch.randn(4, 3), torch.rand(4, 5), None)
    assert torch.allclose(eft.ray_bundle_to_ray_points(rb), eo.ray_bundle_to_ray_points(rb))


def test_no_cpu_path():
    from oracle import eft_oracle as eo
    from sparsefusion_b200.eft import EpipolarFeatureTransformer
    m = EpipolarFeatureTransformer(use_r=True, encoder='resnet18', return_features=True, remove_unused_layers=False)
    images, cams, rb = eo.scene_inputs(image=64, n_rays=4)
    with pytest.raises(RuntimeError):
        m.encode(cams, images)
    with pytest.raises(RuntimeError):
        m.t1(torch.zeros(2, 3, 606))
