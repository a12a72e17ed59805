"""Workloads for ncu: `ncu --profile-from-start off ... python tools/profile_targets.py unet|render|plms`.
Everything before torch.cuda.profiler.start() is warm-up and is not profiled."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparsefusion_b200 import ops  # noqa: E402


def unet(nb=1):
    from sparsefusion_b200.imagen_pytorch import Unet
    u = Unet(channels=4, dim=256, dim_mults=(1, 2, 4, 4), num_resnet_blocks=(2, 2, 2, 2), layer_attns=(False, False, False, True),
             layer_cross_attns=(False,) * 4, cond_images_channels=256, attn_pool_text=False, cond_on_z=False, conditional_embed_dim=None).cuda()
    torch.nn.init.normal_(u.get_parameter('final_conv.weight'), std=0.02)
    x, cond, t = torch.randn(nb, 4, 32, 32, device='cuda'), torch.randn(nb, 256, 32, 32, device='cuda'), torch.full((nb,), 0.3, device='cuda')
    for _ in range(3):
        u.forward(x, t, cond_images=cond)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    u.forward(x, t, cond_images=cond)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def render():
    from sparsefusion_b200.network_grid import NeRFNetwork, get_default_torch_ngp_opt
    opt = get_default_torch_ngp_opt()
    net = NeRFNetwork(opt).cuda().train()
    net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    N = 128 * 128
    o = torch.tensor([0.0, 1.3, 4.8], device='cuda').expand(N, 3).contiguous()
    ys, xs = torch.meshgrid(torch.linspace(1, -1, 128, device='cuda'), torch.linspace(1, -1, 128, device='cuda'), indexing='ij')
    d = torch.stack([xs / 4, ys / 4 - 0.27, -torch.ones_like(xs)], dim=-1).reshape(N, 3).contiguous()
    kw = dict(staged=False, perturb=True, bg_color=0, ambient_ratio=1.0, shading='albedo', force_all_rays=True, **vars(opt))

    def fb():
        net.zero_grad(set_to_none=True)
        r = net.render(o[None], d[None], **kw)
        (r['image'].mean() + r['weights_sum'].mean()).backward()
    for _ in range(3):
        fb()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fb()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == '__main__':
    what = sys.argv[1]
    if what == 'unet':
        unet(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    elif what == 'render':
        render()
