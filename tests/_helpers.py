"""helpers shared by the GPU tests"""
import torch


def device_level_scales(geo):
    """per-level grid scales exactly as the device computes them (exp2f on the GPU; gridencoder.cu:125)"""
    from sparsefusion_b200 import _lib as lib
    s = torch.empty(geo['L'], device='cuda')
    lib.call('sfb_grid_level_scales', geo['L'], float(geo['S']), geo['H'], lib.fptr(s), lib.stream())
    return s.cpu().numpy()
