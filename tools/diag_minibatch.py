"""diagnostic: per-view finiteness / agreement of the batched VAE + PLMS path (small configuration)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_minibatch_gpu import _build, _fixed_target

dev = torch.device('cuda', 0)
res = {}
for V in (1, 2, 4):
    d = _build(0, 1, dev, V)
    seen = {}
    def hook(views, pred_img, seen=seen):
        for j, v in enumerate(views):
            seen[v] = pred_img[j].clone()
        return _fixed_target(views, pred_img)
    d.pred_img_hook = hook
    orig_sample = d.sampler.sample
    def sample(latents, **kw):
        out = orig_sample(latents, **kw)
        print(f'  V={V}: latents finite {torch.isfinite(latents).all().item()} absmax {latents.abs().max().item():.3f}; pred_x0 finite per row '
              f'{[torch.isfinite(out[0][i]).all().item() for i in range(out[0].shape[0])]} absmax {out[0].abs().max().item():.3f}')
        return out
    d.sampler.sample = sample
    d.minibatch_step(1500, max_thres=0.05)
    torch.cuda.synchronize()
    for v, img in seen.items():
        print(f'  V={V} view {v}: pred_img finite {torch.isfinite(img).all().item()} min {img.min().item():.4f} max {img.max().item():.4f} mean {img.mean().item():.4f}')
    res[V] = seen
for v in res[4]:
    for V in (1, 2):
        if v in res[V]:
            print(f'view {v}: batch {V} vs batch 4 max abs diff {(res[V][v] - res[4][v]).abs().max().item():.3e}')
