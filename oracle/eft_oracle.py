"""oracle/eft_oracle.py -- test infrastructure for the Epipolar Feature Transformer (sparsefusion/eft.py).

TEST INFRASTRUCTURE ONLY.  The EFT's python needs pytorch3d (absent here and on the GPU box) only for two camera methods and the RayBundle
container; everything else in sparsefusion/eft.py is plain torch.  So the golden vectors under tests/golden/eft.npz are outputs of the REFERENCE's own
EpipolarFeatureTransformer class, imported unmodified in the build container (oracle/gen_golden.py `eft`) with
  * `pytorch3d.renderer.RayBundle` / `ray_bundle_to_ray_points` replaced by the two-line equivalents below, and
  * cameras given as `NdcCameras`, a restatement of the two PerspectiveCameras methods the EFT calls (eft.py:241 transform_points_ndc,
    eft.py:291 get_camera_center) from pytorch3d's documented conventions (row vectors: x_view = x_world R + T; NDC projection
    x_ndc = f x/z + p) -- pytorch3d itself is not available to check this against: PARITY UNPINNED at the camera boundary, pinned for the network.
`make_params` gives deterministic weights for any module with the EFT's state_dict layout (BatchNorm statistics included).
"""
from __future__ import annotations

import collections
import sys
import types
from typing import Dict

import numpy as np
import torch

RayBundle = collections.namedtuple('RayBundle', 'origins directions lengths xys')


def ray_bundle_to_ray_points(rb):
    return rb.origins[..., None, :] + rb.lengths[..., :, None] * rb.directions[..., None, :]


class NdcCameras:
    """the slice of pytorch3d.renderer.PerspectiveCameras (in_ndc=True) the EFT touches"""

    def __init__(self, R, T, focal_length, principal_point):
        self.R, self.T, self.focal_length, self.principal_point = R, T, focal_length, principal_point

    def __len__(self):
        return self.R.shape[0]

    def get_camera_center(self):
        return -torch.bmm(self.T[:, None, :], self.R.transpose(1, 2))[:, 0, :]

    def transform_points_ndc(self, pts):
        v = torch.bmm(pts.expand(len(self), -1, -1), self.R) + self.T[:, None, :]
        z = v[..., 2:3]
        return torch.cat((self.focal_length[:, None, :] * v[..., :2] / z + self.principal_point[:, None, :], 1.0 / z), dim=-1)


def look_at_cameras(n: int, radius: float = 4.0, elevation_deg: float = 20.0, focal: float = 3.0, device='cpu') -> NdcCameras:
    """n cameras on a circle looking at the origin, pytorch3d axes (+X left, +Y up, +Z forward)"""
    el = np.radians(elevation_deg)
    Rs, Ts = [], []
    for i in range(n):
        az = 2 * np.pi * i / max(n, 1) * 0.35 + 0.3
        c = np.array([radius * np.cos(el) * np.sin(az), radius * np.sin(el), radius * np.cos(el) * np.cos(az)])
        fwd = -c / np.linalg.norm(c)
        left = np.cross(np.array([0.0, 1.0, 0.0]), fwd)
        left /= np.linalg.norm(left)
        up = np.cross(fwd, left)
        M = np.stack([left, up, fwd])               # rows: camera axes in world coordinates
        Rs.append(M.T)                               # x_view = (x - c) M^T = x R + T
        Ts.append(-c @ M.T)
    f32 = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device=device)
    return NdcCameras(f32(Rs), f32(Ts), f32([[focal, focal]] * n), f32([[0.0, 0.0]] * n))


def make_params(shapes: Dict[str, tuple], seed: int = 0) -> Dict[str, torch.Tensor]:
    """deterministic values for a state_dict layout {key: shape}: fan-in-scaled weights, gains around 1, small biases, sane BatchNorm statistics"""
    rng = np.random.default_rng(seed)
    sd = {}
    for k, shape in shapes.items():
        shape = tuple(shape)
        if k.endswith('num_batches_tracked'):
            sd[k] = torch.zeros((), dtype=torch.int64)
            continue
        if k.endswith('running_var'):
            v = rng.random(shape, dtype=np.float32) + 0.5
        elif k.endswith('running_mean'):
            v = rng.standard_normal(shape, dtype=np.float32) * 0.1
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            v = (rng.random(shape, dtype=np.float32) * 2 - 1) * np.sqrt(3.0 / fan_in)
        elif k.endswith('weight'):                    # BatchNorm / LayerNorm gains
            v = rng.random(shape, dtype=np.float32) * 0.4 + 0.8
        else:
            v = rng.standard_normal(shape, dtype=np.float32) * 0.05
        sd[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return sd


def scene_inputs(n_views=2, image=128, n_rays=48, n_depths=20, seed=3):
    """deterministic inputs of one EFT query: input images [NC,3,H,W] in [0,1], input cameras, a flat ray bundle of a query camera"""
    rng = np.random.default_rng(seed)
    images = torch.from_numpy(rng.random((n_views, 3, image, image), dtype=np.float32))
    cams = look_at_cameras(n_views)
    q = look_at_cameras(5)                               # query: another camera of the same ring
    R, T = q.R[3], q.T[3]
    center = -(T @ R.T)
    xy = torch.from_numpy(rng.random((n_rays, 2), dtype=np.float32) * 1.6 - 0.8)
    d_view = torch.cat((xy / 3.0, torch.ones(n_rays, 1)), dim=-1)       # plane at depth 1 (un-normalised directions, as pytorch3d's raysamplers emit)
    directions = d_view @ R.T
    origins = center[None].expand(n_rays, 3).contiguous()
    lengths = torch.linspace(2.0, 6.0, n_depths)[None].expand(n_rays, n_depths).contiguous()
    return images, cams, RayBundle(origins, directions.contiguous(), lengths, None)


def import_reference_eft(ref_root: str = '/root/reference'):
    """the reference's sparsefusion.eft module with its pytorch3d imports stubbed (BUILD CONTAINER ONLY)"""
    import torchvision
    sys.path.insert(0, ref_root)
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
    for n in ('pytorch3d', 'pytorch3d.renderer.cameras', 'pytorch3d.renderer.implicit', 'pytorch3d.renderer.implicit.utils'):
        stub(n)
    r = stub('pytorch3d.renderer')
    r.RayBundle, r.ray_bundle_to_ray_points = RayBundle, ray_bundle_to_ray_points
    orig = torchvision.models.resnet18
    torchvision.models.resnet18 = lambda pretrained=False, **k: orig(weights=None)       # no download: the weights are overwritten by make_params anyway
    try:
        for _ in range(40):
            try:
                import sparsefusion.eft as eft
                break
            except ModuleNotFoundError as e:
                stub(e.name)
    finally:
        pass
    eft._restore_resnet18 = lambda: setattr(torchvision.models, 'resnet18', orig)
    return eft
