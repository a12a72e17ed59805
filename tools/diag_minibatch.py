"""diagnostic: per-view agreement of every stage of the view-batched fusion term across batch compositions (small configuration)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_minibatch_gpu import _build, _fixed_target
from sparsefusion_b200 import image_glue as glue

dev = torch.device('cuda', 0)
res = {}
for V in (1, 2, 4):
    d = _build(0, 1, dev, V)
    rec = {}
    cur = {}
    orig_up = glue.upsample2x_render
    ups = []
    def up_hook(img, ws, h, w, ups=ups):
        out = orig_up(img, ws, h, w)
        ups.append(out.clone())
        return out
    glue.upsample2x_render = up_hook
    orig_encode = d.vae.encode
    def enc(x, cur=cur):
        p = orig_encode(x)
        cur['lat'] = p.mode().clone()
        return p
    d.vae.encode = enc
    orig_sample = d.sampler.sample
    def sample(latents, cur=cur, **kw):
        out = orig_sample(latents, **kw)
        cur['x0'] = out[0].clone()
        return out
    d.sampler.sample = sample
    def hook(views, pred_img, rec=rec, cur=cur, ups=ups):
        for j, v in enumerate(views):
            rec[v] = dict(up=ups[j], lat=cur['lat'][j], x0=cur['x0'][j], img=pred_img[j].clone())
        return _fixed_target(views, pred_img)
    d.pred_img_hook = hook
    d.minibatch_step(1500, max_thres=0.05)
    torch.cuda.synchronize()
    glue.upsample2x_render = orig_up
    print(f'V={V}: views {d.last["views"]}')
    res[V] = rec
rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30)).item()
for v in res[4]:
    for V in (1, 2):
        if v in res[V]:
            a, b = res[V][v], res[4][v]
            print(f'view {v}: batch {V} vs batch 4: render-up rel {rel(a["up"], b["up"]):.2e}  latents {rel(a["lat"], b["lat"]):.2e}  pred_x0 {rel(a["x0"], b["x0"]):.2e}  '
                  f'pred_img rel {rel(a["img"], b["img"]):.2e} maxabs {(a["img"] - b["img"]).abs().max().item():.2e}')
