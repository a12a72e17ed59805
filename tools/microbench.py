"""Per-kernel timings on the GPU box (CUDA events, L2 flushed between iterations) -> gpurun_out/microbench.json.
Used to pick tile/split heuristics and to fill DESIGN.md's per-kernel roofline table; not the headline bench."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sparsefusion_b200 import ops  # noqa: E402

FLUSH = None


def timeit(fn, iters=10, warmup=3, flush=True):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            FLUSH.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def conv_cases():
    # (name, NB scale, H, Cin, Cout, K, stride)
    base = [('init3', 32, 260, 128, 3, 1), ('init7', 32, 260, 64, 7, 1), ('init15', 32, 260, 64, 15, 1), ('d0.conv3', 32, 256, 256, 3, 1),
            ('d0.down4s2', 32, 256, 256, 4, 2), ('d1.conv3', 16, 256, 256, 3, 1), ('d2.conv3', 8, 512, 512, 3, 1),
            ('d3.conv3', 4, 1024, 1024, 3, 1), ('u0.conv3', 4, 2048, 1024, 3, 1), ('u0.res1x1', 4, 2048, 1024, 1, 1),
            ('u1.conv3', 8, 1536, 1024, 3, 1), ('u2.conv3', 16, 768, 512, 3, 1), ('u3.conv3', 32, 512, 256, 3, 1),
            ('ff1x1', 4, 1024, 2048, 1, 1), ('ps1x1', 4, 1024, 4096, 1, 1), ('final', 32, 256, 4, 3, 1)]
    return base


def unet_bench(out):
    import numpy as np
    from sparsefusion_b200 import _lib
    if 'nopdl' in sys.argv:
        _lib.call('sfb_set_pdl', 0)
    if 'nofuse' in sys.argv:
        _lib.call('sfb_set_fusion', 0)
    for a in sys.argv:
        if a.startswith('fuse='):
            _lib.call('sfb_set_fusion', int(a[5:], 0))
    from sparsefusion_b200.imagen_pytorch import Unet, UnetGraph
    unet = Unet(channels=4, dim=256, dim_mults=(1, 2, 4, 4), num_resnet_blocks=(2, 2, 2, 2), layer_attns=(False, False, False, True),
                layer_cross_attns=(False,) * 4, cond_images_channels=256, attn_pool_text=False, cond_on_z=False, conditional_embed_dim=None).cuda()
    torch.nn.init.normal_(unet.get_parameter('final_conv.weight'), std=0.02)
    out['unet'] = {}
    for mode in (('tf32x3',) if 'x3only' in sys.argv else ('tf32x3', 'tf32')):
        ops.set_precision(mode)
        unet.prepare()
        for nb in ((16, 32) if 'nb32' in sys.argv else (1, 8, 16) if 'nb16' in sys.argv else (1, 8)):
            x, cond, t = torch.randn(nb, 4, 32, 32, device='cuda'), torch.randn(nb, 256, 32, 32, device='cuda'), torch.full((nb,), 0.3, device='cuda')
            eager = timeit(lambda: unet.forward(x, t, cond_images=cond), iters=5, warmup=2, flush=False)
            runner = UnetGraph(unet)
            tfeat = unet.precompute_time(t)
            graph = timeit(lambda: runner(x, t, cond, new_cond=False, time_features=tfeat), iters=20, warmup=3, flush=True)
            rec = dict(eager_ms=round(eager, 3), graph_ms=round(graph, 3), tflops=round(62.83e-3 * nb / graph, 1), weight_gbs=round(1602.7 / graph, 1))
            out['unet'][f'{mode}_nb{nb}'] = rec
            print('unet', mode, nb, rec, flush=True)
    ops.set_precision('tf32x3')


def unet_trace(out):
    """in-situ chain cost of every UNet kernel inside the replayed graph (sfb_trace_begin): stamp[i+1] - stamp[i]"""
    global FLUSH
    import ctypes
    import collections
    from sparsefusion_b200 import _lib
    if 'nopdl' in sys.argv:
        _lib.call('sfb_set_pdl', 0)
    if 'nocarve' in sys.argv:
        _lib.call('sfb_set_pdl', 3)
    if 'padsmem' in sys.argv:
        _lib.call('sfb_set_pdl', 5)
    for a in sys.argv:
        if a.startswith('fuse='):
            _lib.call('sfb_set_fusion', int(a[5:], 0))
    from sparsefusion_b200.imagen_pytorch import Unet, UnetGraph
    unet = Unet(channels=4, dim=256, dim_mults=(1, 2, 4, 4), num_resnet_blocks=(2, 2, 2, 2), layer_attns=(False, False, False, True),
                layer_cross_attns=(False,) * 4, cond_images_channels=256, attn_pool_text=False, cond_on_z=False, conditional_embed_dim=None).cuda()
    torch.nn.init.normal_(unet.get_parameter('final_conv.weight'), std=0.02)
    ops.set_precision('tf32x3')
    unet.parallel_res_conv = 'parallel' in sys.argv      # per-kernel attribution needs a single dependency chain
    unet.prepare()
    hw = 32
    for a in sys.argv:
        if a.startswith('hw='):
            hw = int(a[3:])
    x, cond, t = torch.randn(1, 4, hw, hw, device='cuda'), torch.randn(1, 256, hw, hw, device='cuda'), torch.full((1,), 0.3, device='cuda')
    runner = UnetGraph(unet)
    for _ in range(3):
        runner(x, t, cond)
    torch.cuda.synchronize()
    cap = 4096
    buf = torch.zeros(cap + 1, dtype=torch.int64, device='cuda')
    pbuf = torch.zeros(1 + 8 * 128, dtype=torch.int64, device='cuda')
    _lib.call('sfb_trace_begin', buf.data_ptr(), cap)
    tfeat = unet.precompute_time(t)
    feat = unet.precompute_cond(cond)
    _lib.call('sfb_trace_begin', buf.data_ptr(), cap)     # restart: names of the main evaluation only
    buf.zero_()
    unet.forward(x, None, cond_features=feat, time_features=tfeat)            # names in launch order (same sequence as the captured main graph)
    torch.cuda.synchronize()
    n_eager = int(buf[0])
    nbuf = ctypes.create_string_buffer(1 << 20)
    _lib.load().sfb_trace_names(nbuf, len(nbuf))
    names = nbuf.value.decode().split('\n')[:-1]
    _lib.call('sfb_conv_phase_trace', pbuf.data_ptr(), 128)
    runs = []
    for _ in range(5):
        if FLUSH is None:
            FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
        FLUSH.zero_()
        buf.zero_()
        pbuf.zero_()
        torch.cuda.synchronize()
        runner(x, t, cond, new_cond=False, time_features=tfeat)
        torch.cuda.synchronize()
        k = int(buf[0])
        runs.append(buf[1:1 + k].cpu().numpy().astype('int64'))
    _lib.call('sfb_trace_end')
    _lib.call('sfb_conv_phase_trace', None, 0)
    import numpy as np
    names = names[-int(buf[0]):]                     # the eager pass also recorded precompute_cond's kernels first
    st = runs[-1]
    assert len(st) == len(names) <= n_eager, (len(st), len(names), n_eager)
    iv = np.median(np.stack([np.diff(r) for r in runs[1:]]), axis=0) / 1e3      # us, median over replays
    agg = collections.defaultdict(lambda: [0, 0.0])
    for nm, d in zip(names[:-1], iv):
        agg[nm][0] += 1
        agg[nm][1] += float(d)
    total = float(iv.sum())
    print(f'trace: {len(names)} kernels, first-to-last stamp {total:.1f} us', flush=True)
    for nm, (c, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'  {tt:9.1f} us  {c:4d} x {tt / c:7.2f}  {nm}')
    tag = '_'.join(['trace'] + [a for a in sys.argv[1:] if a in ('nopdl', 'nocarve', 'padsmem')])
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    npc = int(pbuf[0])
    json.dump(dict(names=names, interval_us=[round(float(v), 3) for v in iv], total_us=total, stamps=[int(v) for v in runs[-1]],
                   conv_phase_stamps=pbuf[1:1 + 8 * npc].cpu().numpy().reshape(npc, 8).tolist()), open(os.path.join(ROOT, 'gpurun_out', tag + '.json'), 'w'))


def vae_trace(out):
    """chain cost of every kernel of one VAE decode / encode (eager launches queued behind a GPU spin so the stamps show GPU time, not CPU gaps)"""
    import ctypes
    import collections
    import numpy as np
    from sparsefusion_b200 import _lib
    from sparsefusion_b200.ldm_autoencoder import AutoencoderKL
    vae = AutoencoderKL().cuda().eval()
    img = torch.rand(1, 3, 256, 256, device='cuda')
    z = torch.randn(1, 4, 32, 32, device='cuda')
    cap = 4096
    buf = torch.zeros(cap + 1, dtype=torch.int64, device='cuda')
    for what, fn in (('decode', lambda: vae.decode(z)), ('encode', lambda: vae.encode(img).mode())):
        with torch.no_grad():
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            _lib.call('sfb_trace_begin', buf.data_ptr(), cap)
            buf.zero_()
            torch.cuda._sleep(int(6e7))
            fn()
            torch.cuda.synchronize()
        nbuf = ctypes.create_string_buffer(1 << 20)
        _lib.load().sfb_trace_names(nbuf, len(nbuf))
        names = nbuf.value.decode().split('\n')[:-1]
        _lib.call('sfb_trace_end')
        k = int(buf[0])
        st = buf[1:1 + k].cpu().numpy().astype('int64')
        iv = np.diff(st) / 1e3
        agg = collections.defaultdict(lambda: [0, 0.0])
        for nm, d in zip(names[:-1], iv):
            agg[nm][0] += 1
            agg[nm][1] += float(d)
        print(f'vae {what}: {len(names)} kernels, first-to-last {iv.sum():.1f} us')
        for nm, (c, tt) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f'  {tt:9.1f} us  {c:4d} x {tt / c:8.2f}  {nm}')
        top = sorted(zip(iv, range(len(iv))), reverse=True)[:12]
        print('  longest:', [(round(float(d), 1), names[i]) for d, i in top])
        json.dump(dict(names=names, interval_us=[round(float(v), 2) for v in iv]), open(os.path.join(ROOT, 'gpurun_out', f'trace_vae_{what}.json'), 'w'))


def conv_phases(out):
    """where a tcgen05 conv launch spends its time (sfb_conv_phase_trace): medians over 20 launches, cold L2, arena-style accumulate mode"""
    global FLUSH
    import numpy as np
    from sparsefusion_b200 import _lib
    ops.set_precision('tf32x3')
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    cap = 64
    buf = torch.zeros(1 + 8 * cap, dtype=torch.int64, device='cuda')
    print('phases (us): prologue | dep-wait | first-stage | mainloop | drain | epilogue | total   [grid]')
    rows = []
    for name, h, cin, cout, k, stride in conv_cases():
        x = torch.randn(1, h, h, cin, device='cuda')
        w = ops.pack_conv_weight(torch.randn(cout, cin, k, k, device='cuda') / (cin * k * k) ** 0.5)
        b = torch.zeros(cout, device='cuda')
        pad = (k - 1) // 2 if stride == 1 else 1
        y = ops.conv2d_nhwc(x, w, cout, k, k, stride, pad, bias=b)
        y.zero_()
        _lib.call('sfb_conv_phase_trace', buf.data_ptr(), cap)
        buf.zero_()
        n = 20
        for _ in range(n):
            FLUSH.zero_()
            ops.conv2d_nhwc(x, w, cout, k, k, stride, pad, bias=b, out=y, accumulate=True)
        torch.cuda.synchronize()
        _lib.call('sfb_conv_phase_trace', None, 0)
        st = buf[1:1 + 8 * n].cpu().numpy().reshape(n, 8)[:, :7].astype('float64')
        d = np.median(np.diff(st, axis=1), axis=0) / 1e3
        tot = float(np.median(st[:, 6] - st[:, 0]) / 1e3)
        rec = dict(name=name, hw=h, cin=cin, cout=cout, k=k, prologue=round(float(d[0]), 2), dep_wait=round(float(d[1]), 2), first_stage=round(float(d[2]), 2),
                   mainloop=round(float(d[3]), 2), drain=round(float(d[4]), 2), epilogue=round(float(d[5]), 2), total=round(tot, 2), weight_mb=round(w.numel() * 4 / 1e6, 2))
        rows.append(rec)
        print(f"{name:12s} {d[0]:6.2f} | {d[1]:6.2f} | {d[2]:6.2f} | {d[3]:6.2f} | {d[4]:6.2f} | {d[5]:6.2f} | {tot:6.2f}   W {rec['weight_mb']:6.2f} MB", flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'conv_phases.json'), 'w'), indent=1)


def vae_bench(out):
    from sparsefusion_b200.ldm_autoencoder import AutoencoderKL
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    vae = AutoencoderKL().cuda().eval()
    img = torch.rand(1, 3, 256, 256, device='cuda')
    z = torch.randn(1, 4, 32, 32, device='cuda')
    with torch.no_grad():
        out['vae_sm100'] = dict(encode_ms=round(timeit(lambda: vae.encode(img).mode(), iters=10, flush=False), 3),
                                decode_ms=round(timeit(lambda: vae.decode(z), iters=10, flush=False), 3))
        print('vae sm100', out['vae_sm100'], flush=True)
    from oracle import vae_oracle as vo                     # eager torch / cuDNN (TF32) restatement as the comparison
    tv = vo.TorchVAE(vae.state_dict()).to('cuda')
    with torch.no_grad():
        enc = timeit(lambda: tv.encode(img).mode(), iters=10, flush=False)
        dec = timeit(lambda: tv.decode(z), iters=10, flush=False)
        enc_cl = dec_cl = float('nan')
    out['vae'] = dict(encode_ms=round(enc, 3), decode_ms=round(dec, 3), encode_cl_ms=round(enc_cl, 3), decode_cl_ms=round(dec_cl, 3))
    print('vae', out['vae'], flush=True)


def render_bench(out):
    import numpy as np
    from sparsefusion_b200 import _lib
    for a in sys.argv:
        if a.startswith('fuse='):
            _lib.call('sfb_set_fusion', int(a[5:], 0))
    from sparsefusion_b200.network_grid import NeRFNetwork, get_default_torch_ngp_opt
    opt = get_default_torch_ngp_opt()
    net = NeRFNetwork(opt).cuda().train()
    net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    N = 128 * 128
    o = torch.tensor([0.0, 1.3, 4.8], device='cuda').expand(N, 3).contiguous()
    ys, xs = torch.meshgrid(torch.linspace(1, -1, 128, device='cuda'), torch.linspace(1, -1, 128, device='cuda'), indexing='ij')
    d = torch.stack([xs / 4, ys / 4 - 0.27, -torch.ones_like(xs)], dim=-1).reshape(N, 3).contiguous()
    kw = dict(staged=False, perturb=True, bg_color=0, ambient_ratio=1.0, shading='albedo', force_all_rays=True, **vars(opt))
    def fwd():
        with torch.no_grad():
            net.render(o[None], d[None], **kw)
    def fwd_bwd():
        net.zero_grad(set_to_none=True)
        r = net.render(o[None], d[None], **kw)
        (r['image'].mean() + r['weights_sum'].mean()).backward()
    out['render'] = dict(rays=N, fwd_ms=round(timeit(fwd, iters=10), 3), fwd_bwd_ms=round(timeit(fwd_bwd, iters=10), 3))
    print('render', out['render'], flush=True)


def main():
    out = {'device': torch.cuda.get_device_name(0), 'conv': [], 'grid': {}}
    if 'unet' in sys.argv or len(sys.argv) == 1:
        unet_bench(out)
    if 'render' in sys.argv or len(sys.argv) == 1:
        render_bench(out)
    if 'trace' in sys.argv:
        unet_trace(out)
    if 'phases' in sys.argv:
        conv_phases(out)
    if 'vaetrace' in sys.argv:
        vae_trace(out)
    if 'vae' in sys.argv or len(sys.argv) == 1:
        vae_bench(out)
    if len(sys.argv) > 1 and 'conv' not in sys.argv:
        json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'microbench.json'), 'w'), indent=1)
        return
    for nb in (1, 8):
        for name, h, cin, cout, k, stride in conv_cases():
            x = ops.round_tf32(torch.randn(nb, h, h, cin, device='cuda'))
            w = ops.pack_conv_weight(torch.randn(cout, cin, k, k, device='cuda') / (cin * k * k) ** 0.5)
            b = torch.zeros(cout, device='cuda')
            pad = (k - 1) // 2 if stride == 1 else 1
            y = ops.conv2d_nhwc(x, w, cout, k, k, stride, pad, bias=b)
            ms = timeit(lambda: ops.conv2d_nhwc(x, w, cout, k, k, stride, pad, bias=b, out=y))
            ho = y.shape[1]
            flops = 2.0 * nb * ho * ho * cout * cin * k * k
            wbytes = w.numel() * 4
            rec = dict(name=name, nb=nb, hw=h, cin=cin, cout=cout, k=k, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1),
                       weight_gbs=round(wbytes / ms / 1e6, 1))
            print(rec, flush=True)
            out['conv'].append(rec)
    # grid encoder operators at the render's sizes
    from sparsefusion_b200.gridencoder import GridEncoder
    enc = GridEncoder(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=16, desired_resolution=8192, gridtype='tiled').cuda()
    enc.embeddings.data.uniform_(-0.5, 0.5)
    for B in (1 << 20, 1 << 21):
        x = (torch.rand(B, 3, device='cuda') * 2 - 1) * 4
        ms_f = timeit(lambda: enc(x, bound=4).sum() if False else enc(x, bound=4))
        y = enc(x, bound=4)
        g = torch.randn_like(y)
        def fb():
            enc.embeddings.grad = None
            y2 = enc(x, bound=4)
            y2.backward(g)
        ms_fb = timeit(fb)
        out['grid'][str(B)] = dict(fwd_ms=round(ms_f, 4), fwd_bwd_ms=round(ms_fb, 4), fwd_GBps_alg=round(B * 140 / ms_f / 1e6, 1))
        print('grid', B, out['grid'][str(B)], flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'microbench.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
