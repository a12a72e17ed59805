"""oracle/lpips_oracle.py -- CPU restatement of the LPIPS-VGG perceptual distance the distillation loop adds to the fusion loss
(sparsefusion/distillation.py:161, :312-314 through external/external_utils.py:11-49: `lpips.LPIPS(net='vgg')(2*pred-1, 2*target-1)`).

TEST INFRASTRUCTURE ONLY (tests/, bench.py's CPU arm, __graft_entry__.smoke).

**Parity unpinned.**  The algorithm lives in the third-party `lpips` package (Zhang et al., CVPR 2018; the reference's ENVIRONMENT.md installs it
unpinned with pip, 0.1.4 at the time) which is NOT under /root/reference and not installed here, and its arithmetic depends on pretrained
weights (torchvision VGG16 + the learned `lin` layers) that cannot be fetched.  This file restates the published forward:

    x  -> (x - shift) / scale                       shift = (-.030, -.088, -.188), scale = (.458, .448, .450)          (ScalingLayer)
    f_k = VGG16 features after relu1_2, relu2_2, relu3_3, relu4_3, relu5_3   (64, 128, 256, 512, 512 channels; 3x3 convs, 2x2 max-pools)
    n_k = f_k / (||f_k||_2 over channels + 1e-10)                                                                         (normalize_tensor)
    d   = sum_k  mean_{h,w}  sum_c  w_k[c] * (n_k(x0) - n_k(x1))[c]^2         w_k >= 0, the 1x1 `lin` layers (bias-free; dropout is identity in eval)

with SEEDED RANDOM weights of the right shapes (He-scaled convolutions so activations stay O(1) through 13 layers, non-negative lin weights).
What the tests pin is therefore: the product kernels == this restatement (value and gradient w.r.t. x0) for identical weights, and the BACKBONE
half of the restatement == torchvision 0.26's `vgg16().features` sliced where lpips slices it, for identical (random) weights
(tests/test_oracle_lpips.py::test_backbone_taps_match_torchvision_vgg16); what they cannot pin is the `lin` heads / normalisation against the
package's code, or equality with the pretrained network's numbers.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512]   # torchvision vgg16.features up to relu5_3
TAPS = (1, 3, 6, 9, 12)                     # index (0-based) of the conv after whose ReLU a feature map is taken: relu1_2, 2_2, 3_3, 4_3, 5_3
CHNS = (64, 128, 256, 512, 512)
SHIFT = (-.030, -.088, -.188)
SCALE = (.458, .448, .450)


def make_params(seed: int = 0) -> Dict[str, torch.Tensor]:
    """{'conv{i}.weight' [Cout,Cin,3,3], 'conv{i}.bias' [Cout] for the 13 convolutions, 'lin{k}.weight' [1,C_k,1,1] for the 5 heads}"""
    rng = np.random.default_rng(seed)
    p: Dict[str, torch.Tensor] = {}
    cin, i = 3, 0
    for v in CFG:
        if v == 'M':
            continue
        w = rng.standard_normal((v, cin, 3, 3)).astype(np.float32) * np.sqrt(2.0 / (cin * 9))
        p[f'conv{i}.weight'] = torch.from_numpy(w)
        p[f'conv{i}.bias'] = torch.from_numpy((rng.standard_normal(v) * 0.05).astype(np.float32))
        cin, i = v, i + 1
    for k, c in enumerate(CHNS):
        p[f'lin{k}.weight'] = torch.from_numpy((rng.random((1, c, 1, 1)).astype(np.float32)) / c * 4)
    return p


def features(p: Dict[str, torch.Tensor], x: torch.Tensor) -> List[torch.Tensor]:
    """x [B,3,H,W] in [-1,1] -> the five tapped feature maps"""
    shift = torch.tensor(SHIFT, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    scale = torch.tensor(SCALE, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    h = (x - shift) / scale
    outs, i = [], 0
    for v in CFG:
        if v == 'M':
            h = F.max_pool2d(h, 2, 2)
            continue
        h = F.relu(F.conv2d(h, p[f'conv{i}.weight'].to(x.dtype), p[f'conv{i}.bias'].to(x.dtype), padding=1))
        if i in TAPS:
            outs.append(h)
        i += 1
    return outs


def lpips(p: Dict[str, torch.Tensor], x0: torch.Tensor, x1: torch.Tensor) -> torch.Tensor:
    """[B,1,1,1] distance of images already scaled to [-1,1] (the reference passes normalize=True, i.e. 2*img-1, external_utils.py:37-39)"""
    f0, f1 = features(p, x0), features(p, x1)
    total = 0
    for k in range(5):
        n0 = f0[k] / (f0[k].pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-10)
        n1 = f1[k] / (f1[k].pow(2).sum(dim=1, keepdim=True).sqrt() + 1e-10)
        d = (n0 - n1) ** 2
        total = total + (d * p[f'lin{k}.weight'].to(x0.dtype)).sum(dim=1, keepdim=True).mean(dim=(2, 3), keepdim=True)
    return total


class PerceptualLoss:
    """external/external_utils.py:11-49: inputs in [0,1], normalize=True -> [-1,1]; returns [B,1,1,1]"""

    def __init__(self, params: Dict[str, torch.Tensor]):
        self.p = params

    def __call__(self, pred, target, normalize=True):
        if normalize:
            target = 2 * target - 1
            pred = 2 * pred - 1
        return lpips(self.p, pred.float(), target.float())
