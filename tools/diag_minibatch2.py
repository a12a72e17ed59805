"""bisect: where does the view-batched step's decoded image go wrong (V = 1, small configuration)?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_minibatch_gpu import _build, _fixed_target

dev = torch.device('cuda', 0)


def run(tag, V=1, skip_photo=False, graph=True, poke=None):
    d = _build(0, 1, dev, V)
    d.sampler.use_cuda_graph = graph
    eng_sum = lambda: sum(float(w.double().sum()) for w in d.vae._sm100.W.values()) if d.vae._sm100 is not None else None
    zfix = torch.randn(1, 4, 16, 16, device=dev, generator=torch.Generator(device=dev).manual_seed(9))
    with torch.no_grad():
        ref_dec = d.vae.decode(zfix).clone()
    s0 = eng_sum()
    info = {}
    orig_decode = d.vae.decode

    def dec(z):
        info['z_finite'] = bool(torch.isfinite(z).all())
        info['z'] = z.clone()
        out = orig_decode(z)
        info['out_finite'] = bool(torch.isfinite(out).all())
        return out
    d.vae.decode = dec
    d.pred_img_hook = _fixed_target
    if skip_photo:
        from sparsefusion_b200 import image_glue as glue
        orig = glue.photometric_loss
        d_itr = 1501      # itr % R == r always true for R = 1: emulate the skip by zeroing the gradient instead
    d.minibatch_step(1500, max_thres=0.05)
    torch.cuda.synchronize()
    s1 = eng_sum()
    with torch.no_grad():
        again = orig_decode(info['z'])
        ref2 = orig_decode(zfix)
    print(f'{tag:34s} decode input finite {info["z_finite"]} output finite {info["out_finite"]}; same input decoded again finite {bool(torch.isfinite(again).all())}; '
          f'fixed-z decode drift {(ref2 - ref_dec).abs().max().item():.2e}; packed-weight checksum {s0:.6f} -> {s1:.6f}', flush=True)


run('V=1 default')
run('V=1 default (again)')
run('V=1 no cuda graph', graph=False)
run('V=2 default', V=2)
run('V=4 default', V=4)
