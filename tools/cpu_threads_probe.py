"""How many intra-op threads does the CPU port of a UNet evaluation want on this host? (informs bench.py's cpu_baseline / --impl reference)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import unet_oracle as uo
cfg = uo.FULL
sd = uo.make_params(cfg, seed=0)
x, c = torch.randn(1, 4, 32, 32), torch.randn(1, 256, 32, 32)
ls = uo.alpha_cosine_log_snr(torch.tensor([0.3]))
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for n in (128, 64, 32, 16, 8):
    torch.set_num_threads(n)
    with torch.no_grad():
        uo.unet_forward(sd, cfg, x, ls, c)
        t0 = time.perf_counter(); uo.unet_forward(sd, cfg, x, ls, c); dt = time.perf_counter() - t0
    print(f'threads {n:4d}: {dt:.2f} s per UNet evaluation', flush=True)
    if dt > 20 and n <= 32:
        break
