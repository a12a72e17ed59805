"""one view-batched step (small configuration) -- the workload for compute-sanitizer"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from test_minibatch_gpu import _build, _fixed_target
V = int(sys.argv[1]) if len(sys.argv) > 1 else 2
d = _build(0, 1, torch.device('cuda', 0), V)
if 'nograph' in sys.argv:
    d.sampler.use_cuda_graph = False
seen = {}
def hook(views, pred_img):
    for j, v in enumerate(views):
        seen[v] = bool(torch.isfinite(pred_img[j]).all())
    return _fixed_target(views, pred_img)
d.pred_img_hook = hook
d.minibatch_step(1500, max_thres=0.05)
torch.cuda.synchronize()
print('finite per view:', seen, 'grad finite', bool(torch.isfinite(d.optimizer.grad).all()))
