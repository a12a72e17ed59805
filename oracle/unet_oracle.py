"""oracle/unet_oracle.py -- CPU restatement of the reference's VLDM UNet forward and PLMS sampler.

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs as the checker and the CPU baseline.  Nothing under
``sparsefusion_b200/`` imports it.

What it restates (all citations relative to the reference repo):

* ``Unet.forward`` / ``forward_with_cond_scale``  external/imagen_pytorch.py:1456-1671, in the
  live configuration of utils/load_model.py:58-69 after ``DDPM.__init__`` re-instantiates it with
  ``cond_on_z=False, lowres_cond=False`` (sparsefusion/vldm.py:165-171): no text conditioning, no
  attention pooling, ``layer_cross_attns`` all False, ``memory_efficient=False``,
  ``init_cross_embed=True``, ``pixel_shuffle_upsample=True``, ``scale_skip_connection=True``,
  ``final_resnet_block=True``, gca on the "extra" resnet blocks.
* blocks: ``Block`` :641-662, ``ResnetBlock`` :664-729, ``CrossAttention`` :731-805,
  ``Attention`` :480-566, ``GlobalContext`` :916-941, ``ChanFeedForward`` :953-961,
  ``TransformerBlock`` :963-988, ``CrossEmbedLayer`` :1017-1042, ``PixelShuffleUpsample`` :578-606,
  ``LayerNorm``/``ChanLayerNorm`` :301-329, ``LearnedSinusoidalPosEmb`` :624-639.
* ``GaussianDiffusionContinuousTimes``  :201-297 and ``alpha_cosine_log_snr`` :194-196.
* ``PLMSSampler``  external/plms.py:13-214, with every ``randn_like`` draw injectable.

The restatement is a plain function of a ``{reference state_dict key: tensor}`` mapping, so it
runs on any box (the GPU box has no /root/reference).  It is pinned against the reference's own
classes by ``oracle/gen_golden.py`` (run in the build container, where the reference imports):
``tests/golden/unet_*.npz`` hold the reference's outputs and ``tests/test_oracle_unet.py`` checks
this file reproduces them.

It works in whatever dtype the parameters are given in (fp32 for the CPU baseline, fp64 to
calibrate the TF32 error budget of the CUDA path).
"""
from __future__ import annotations

import math
import zlib
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# configuration (utils/load_model.py:58-69 + imagen_pytorch.py:1082-1126 defaults)
# ---------------------------------------------------------------------------------------------
@dataclass
class UnetConfig:
    dim: int = 256
    dim_mults: Tuple[int, ...] = (1, 2, 4, 4)
    num_resnet_blocks: Tuple[int, ...] = (2, 2, 2, 2)
    layer_attns: Tuple[bool, ...] = (False, False, False, True)
    channels: int = 4
    cond_images_channels: int = 256
    attn_dim_head: int = 64
    attn_heads: int = 8
    ff_mult: float = 2.0
    num_time_tokens: int = 2
    learned_sinu_pos_emb_dim: int = 16
    resnet_groups: int = 8
    init_cross_embed_kernel_sizes: Tuple[int, ...] = (3, 7, 15)
    image_size: int = 32

    @property
    def cond_dim(self) -> int:
        return self.dim

    @property
    def time_cond_dim(self) -> int:
        return self.dim * 4

    def reference_kwargs(self) -> dict:
        """kwargs for the reference ``Unet`` constructor equivalent to this config."""
        return dict(channels=self.channels, dim=self.dim, dim_mults=self.dim_mults,
                    num_resnet_blocks=self.num_resnet_blocks, layer_attns=self.layer_attns,
                    layer_cross_attns=tuple(False for _ in self.dim_mults),
                    cond_images_channels=self.cond_images_channels, attn_pool_text=False,
                    attn_dim_head=self.attn_dim_head, attn_heads=self.attn_heads, ff_mult=self.ff_mult,
                    cond_on_z=False, conditional_embed_dim=None)


CROSS_DIM_HEAD, CROSS_HEADS = 64, 8

FULL = UnetConfig()
SMALL = UnetConfig(dim=32, cond_images_channels=12, attn_dim_head=16, attn_heads=2, image_size=16)


def param_shapes(cfg: UnetConfig) -> Dict[str, Tuple[int, ...]]:
    """Every state_dict key of the live reference Unet with its shape (477 tensors for FULL)."""
    s: Dict[str, Tuple[int, ...]] = {}
    dim, cd, td = cfg.dim, cfg.cond_dim, cfg.time_cond_dim
    inner = cfg.attn_dim_head * cfg.attn_heads
    dh = cfg.attn_dim_head
    s['null_conditional_embed'] = (1, 256, cd)
    s['null_conditional_hidden'] = (1, td)
    cin = cfg.channels + cfg.cond_images_channels
    ks = sorted(cfg.init_cross_embed_kernel_sizes)
    dim_scales = [int(dim / (2 ** i)) for i in range(1, len(ks))]
    dim_scales = [*dim_scales, dim - sum(dim_scales)]
    for i, (k, dsc) in enumerate(zip(ks, dim_scales)):
        s[f'init_conv.convs.{i}.weight'] = (dsc, cin, k, k)
        s[f'init_conv.convs.{i}.bias'] = (dsc,)
    s['to_time_hiddens.0.weights'] = (cfg.learned_sinu_pos_emb_dim // 2,)
    s['to_time_hiddens.1.weight'] = (td, cfg.learned_sinu_pos_emb_dim + 1)
    s['to_time_hiddens.1.bias'] = (td,)
    s['to_time_cond.0.weight'] = (td, td)
    s['to_time_cond.0.bias'] = (td,)
    s['to_time_tokens.0.weight'] = (cd * cfg.num_time_tokens, td)
    s['to_time_tokens.0.bias'] = (cd * cfg.num_time_tokens,)
    s['norm_cond.weight'] = (cd,)
    s['norm_cond.bias'] = (cd,)

    def resnet(p, din, dout, gca, cross):
        s[f'{p}.time_mlp.1.weight'] = (dout * 2, td)
        s[f'{p}.time_mlp.1.bias'] = (dout * 2,)
        if cross:
            # mid blocks are built with ResnetBlock(...) directly (imagen_pytorch.py:1336-1338), not
            # resnet_klass, so CrossAttention keeps its own defaults dim_head=64, heads=8 (:737-738)
            s[f'{p}.cross_attn.fn.null_kv'] = (2, CROSS_DIM_HEAD)
            s[f'{p}.cross_attn.fn.norm.g'] = (dout,)
            s[f'{p}.cross_attn.fn.to_q.weight'] = (CROSS_DIM_HEAD * CROSS_HEADS, dout)
            s[f'{p}.cross_attn.fn.to_kv.weight'] = (CROSS_DIM_HEAD * CROSS_HEADS * 2, cd)
            s[f'{p}.cross_attn.fn.to_out.0.weight'] = (dout, CROSS_DIM_HEAD * CROSS_HEADS)
            s[f'{p}.cross_attn.fn.to_out.1.g'] = (dout,)
        s[f'{p}.block1.groupnorm.weight'] = (din,)
        s[f'{p}.block1.groupnorm.bias'] = (din,)
        s[f'{p}.block1.project.weight'] = (dout, din, 3, 3)
        s[f'{p}.block1.project.bias'] = (dout,)
        s[f'{p}.block2.groupnorm.weight'] = (dout,)
        s[f'{p}.block2.groupnorm.bias'] = (dout,)
        s[f'{p}.block2.project.weight'] = (dout, dout, 3, 3)
        s[f'{p}.block2.project.bias'] = (dout,)
        if gca:
            hid = max(3, dout // 2)
            s[f'{p}.gca.to_k.weight'] = (1, dout, 1, 1)
            s[f'{p}.gca.to_k.bias'] = (1,)
            s[f'{p}.gca.net.0.weight'] = (hid, dout, 1, 1)
            s[f'{p}.gca.net.0.bias'] = (hid,)
            s[f'{p}.gca.net.2.weight'] = (dout, hid, 1, 1)
            s[f'{p}.gca.net.2.bias'] = (dout,)
        if din != dout:
            s[f'{p}.res_conv.weight'] = (dout, din, 1, 1)
            s[f'{p}.res_conv.bias'] = (dout,)

    def attention(p, d, context):
        s[f'{p}.null_kv'] = (2, dh)
        s[f'{p}.norm.g'] = (d,)
        s[f'{p}.to_q.weight'] = (inner, d)
        s[f'{p}.to_kv.weight'] = (dh * 2, d)
        if context:
            s[f'{p}.to_context.0.weight'] = (cd,)
            s[f'{p}.to_context.0.bias'] = (cd,)
            s[f'{p}.to_context.1.weight'] = (dh * 2, cd)
            s[f'{p}.to_context.1.bias'] = (dh * 2,)
        s[f'{p}.to_out.0.weight'] = (d, inner)
        s[f'{p}.to_out.1.g'] = (d,)

    def transformer(p, d):
        attention(f'{p}.layers.0.0.fn', d, True)
        hid = int(d * cfg.ff_mult)
        s[f'{p}.layers.0.1.0.g'] = (1, d, 1, 1)
        s[f'{p}.layers.0.1.1.weight'] = (hid, d, 1, 1)
        s[f'{p}.layers.0.1.3.g'] = (1, hid, 1, 1)
        s[f'{p}.layers.0.1.4.weight'] = (d, hid, 1, 1)

    dims = [dim, *[dim * m for m in cfg.dim_mults]]
    in_out = list(zip(dims[:-1], dims[1:]))
    n = len(in_out)
    for i, (din, dout) in enumerate(in_out):
        last = i == n - 1
        resnet(f'downs.{i}.1', din, din, False, False)
        for j in range(cfg.num_resnet_blocks[i]):
            resnet(f'downs.{i}.2.{j}', din, din, True, False)
        if cfg.layer_attns[i]:
            transformer(f'downs.{i}.3', din)
        if not last:
            s[f'downs.{i}.4.weight'] = (dout, din, 4, 4)
            s[f'downs.{i}.4.bias'] = (dout,)
        else:
            s[f'downs.{i}.4.fns.0.weight'] = (dout, din, 3, 3)
            s[f'downs.{i}.4.fns.0.bias'] = (dout,)
            s[f'downs.{i}.4.fns.1.weight'] = (dout, din, 1, 1)
            s[f'downs.{i}.4.fns.1.bias'] = (dout,)
    mid = dims[-1]
    resnet('mid_block1', mid, mid, False, True)
    attention('mid_attn.fn.fn', mid, False)
    resnet('mid_block2', mid, mid, False, True)
    for i, (din, dout) in enumerate(reversed(in_out)):
        last = i == n - 1
        ri = n - 1 - i
        skip = din  # skip_connect_dims: current_dim == dim_in of the mirrored down stage
        resnet(f'ups.{i}.0', dout + skip, dout, False, False)
        for j in range(cfg.num_resnet_blocks[ri]):
            resnet(f'ups.{i}.1.{j}', dout + skip, dout, True, False)
        if cfg.layer_attns[ri]:
            transformer(f'ups.{i}.2', dout)
        if not last:
            s[f'ups.{i}.3.net.0.weight'] = (din * 4, dout, 1, 1)
            s[f'ups.{i}.3.net.0.bias'] = (din * 4,)
    resnet('final_res_block', dim, dim, True, False)
    s['final_conv.weight'] = (cfg.channels, dim, 3, 3)
    s['final_conv.bias'] = (cfg.channels,)
    return s


def make_params(cfg: UnetConfig, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Deterministic, platform-independent synthetic weights keyed by reference state_dict names.

    Each tensor is drawn from numpy's PCG64 stream seeded by crc32(name) ^ seed (uniform, scaled to
    a fan-in-normalised std), so the GPU box regenerates bit-identical weights without shipping
    1.6 GB.  ``final_conv`` is NOT zero (the reference zero-initialises it, imagen_pytorch.py:1388,
    which would make every parity test vacuous -- SURVEY.md §0.8); gains are 1 +- 0.1.
    """
    out: Dict[str, torch.Tensor] = {}
    for name, shape in param_shapes(cfg).items():
        rng = np.random.default_rng((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0xFFFFFFFF)
        u = rng.random(size=shape, dtype=np.float32) * 2.0 - 1.0  # U(-1,1), std 1/sqrt(3)
        leaf = name.split('.')[-1]
        if leaf in ('g',) or (leaf == 'weight' and ('groupnorm' in name or name.startswith('norm_cond')
                                                    or name.endswith('to_context.0.weight'))):
            t = 1.0 + 0.1 * u
        elif leaf == 'bias':
            t = 0.05 * u
        elif leaf == 'null_kv' or name.startswith('null_conditional'):
            t = u * math.sqrt(3.0)
        elif leaf == 'weights':  # learned sinusoidal frequencies ~ N(0,1) in the reference
            t = u * math.sqrt(3.0)
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            t = u * math.sqrt(3.0) / math.sqrt(fan_in)
        out[name] = torch.from_numpy(np.ascontiguousarray(t)).to(dtype)
    return out


# ---------------------------------------------------------------------------------------------
# diffusion schedule (imagen_pytorch.py:194-297)
# ---------------------------------------------------------------------------------------------
def alpha_cosine_log_snr(t: torch.Tensor, s: float = 0.008) -> torch.Tensor:
    return -torch.log(((torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** -2) - 1).clamp(min=1e-5))


def log_snr_to_alpha_sigma(log_snr):
    return torch.sqrt(torch.sigmoid(log_snr)), torch.sqrt(torch.sigmoid(-log_snr))


def _pad(x, t):
    return t.view(*t.shape, *((1,) * (x.ndim - t.ndim)))


def q_sample(x_start, t, noise):
    log_snr = alpha_cosine_log_snr(t)
    alpha, sigma = log_snr_to_alpha_sigma(_pad(x_start, log_snr))
    return alpha * x_start + sigma * noise, log_snr


def predict_start_from_noise(x_t, t, noise):
    alpha, sigma = log_snr_to_alpha_sigma(_pad(x_t, alpha_cosine_log_snr(t)))
    return (x_t - sigma * noise) / alpha.clamp(min=1e-8)


def q_posterior(x_start, x_t, t, t_next):
    log_snr, log_snr_next = _pad(x_t, alpha_cosine_log_snr(t)), _pad(x_t, alpha_cosine_log_snr(t_next))
    alpha, sigma = log_snr_to_alpha_sigma(log_snr)
    alpha_next, sigma_next = log_snr_to_alpha_sigma(log_snr_next)
    c = -torch.expm1(log_snr - log_snr_next)
    mean = alpha_next * (x_t * (1 - c) / alpha + c * x_start)
    var = (sigma_next ** 2) * c
    return mean, var, torch.log(var.clamp(min=1e-20))


# ---------------------------------------------------------------------------------------------
# blocks
# ---------------------------------------------------------------------------------------------
def _ln(x, g, dim=-1):
    eps = 1e-5 if x.dtype in (torch.float32, torch.float64) else 1e-3
    var = torch.var(x, dim=dim, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=dim, keepdim=True)
    return (x - mean) * (var + eps).rsqrt() * g


class _P:
    """prefix view on the flat parameter dict"""

    def __init__(self, sd, prefix=''):
        self.sd, self.prefix = sd, prefix

    def __getitem__(self, k):
        return self.sd[self.prefix + k]

    def has(self, k):
        return (self.prefix + k) in self.sd

    def sub(self, p):
        return _P(self.sd, self.prefix + p + '.')


def _block(p: _P, x, groups, scale_shift=None):
    x = F.group_norm(x, groups, p['groupnorm.weight'], p['groupnorm.bias'], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    x = F.silu(x)
    return F.conv2d(x, p['project.weight'], p['project.bias'], padding=1)


def _cross_attention(p: _P, x, context, heads):
    b, n, _ = x.shape
    x = _ln(x, p['norm.g'])
    q = F.linear(x, p['to_q.weight'])
    k, v = F.linear(context, p['to_kv.weight']).chunk(2, dim=-1)
    split = lambda t: t.view(b, t.shape[1], heads, -1).transpose(1, 2)
    q, k, v = split(q), split(k), split(v)
    nk, nv = p['null_kv'].unbind(dim=-2)
    nk = nk.view(1, 1, 1, -1).expand(b, heads, 1, -1)
    nv = nv.view(1, 1, 1, -1).expand(b, heads, 1, -1)
    k = torch.cat((nk, k), dim=-2)
    v = torch.cat((nv, v), dim=-2)
    q = q * (q.shape[-1] ** -0.5)
    sim = torch.einsum('bhid,bhjd->bhij', q, k)
    attn = sim.softmax(dim=-1)
    out = torch.einsum('bhij,bhjd->bhid', attn, v)
    out = out.transpose(1, 2).reshape(b, n, -1)
    return _ln(F.linear(out, p['to_out.0.weight']), p['to_out.1.g'])


def _attention(p: _P, x, context, heads):
    """multi-query self attention with null kv and optional context kv (imagen_pytorch.py:511-566)"""
    b, n, _ = x.shape
    x = _ln(x, p['norm.g'])
    q = F.linear(x, p['to_q.weight'])
    k, v = F.linear(x, p['to_kv.weight']).chunk(2, dim=-1)
    q = q.view(b, n, heads, -1).transpose(1, 2)
    q = q * (q.shape[-1] ** -0.5)
    nk, nv = p['null_kv'].unbind(dim=-2)
    k = torch.cat((nk.view(1, 1, -1).expand(b, 1, -1), k), dim=-2)
    v = torch.cat((nv.view(1, 1, -1).expand(b, 1, -1), v), dim=-2)
    if context is not None:
        c = F.layer_norm(context, (context.shape[-1],), p['to_context.0.weight'], p['to_context.0.bias'])
        ck, cv = F.linear(c, p['to_context.1.weight'], p['to_context.1.bias']).chunk(2, dim=-1)
        k = torch.cat((ck, k), dim=-2)
        v = torch.cat((cv, v), dim=-2)
    sim = torch.einsum('bhid,bjd->bhij', q, k)
    attn = sim.softmax(dim=-1)
    out = torch.einsum('bhij,bjd->bhid', attn, v)
    out = out.transpose(1, 2).reshape(b, n, -1)
    return _ln(F.linear(out, p['to_out.0.weight']), p['to_out.1.g'])


def _to_seq(x):
    b, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(b, h * w, c)


def _to_map(x, h, w):
    b, n, c = x.shape
    return x.view(b, h, w, c).permute(0, 3, 1, 2)


def _gca(p: _P, x):
    b, c, h, w = x.shape
    context = F.conv2d(x, p['to_k.weight'], p['to_k.bias']).view(b, 1, h * w)
    out = torch.einsum('bin,bcn->bci', context.softmax(dim=-1), x.view(b, c, h * w)).unsqueeze(-1)
    out = F.silu(F.conv2d(out, p['net.0.weight'], p['net.0.bias']))
    return torch.sigmoid(F.conv2d(out, p['net.2.weight'], p['net.2.bias']))


def _resnet(p: _P, x, t, cond, cfg: UnetConfig, taps=None, name=''):
    scale_shift = None
    if t is not None:
        te = F.linear(F.silu(t), p['time_mlp.1.weight'], p['time_mlp.1.bias'])
        scale_shift = te[:, :, None, None].chunk(2, dim=1)
    h = _block(p.sub('block1'), x, cfg.resnet_groups)
    if p.has('cross_attn.fn.null_kv'):
        hh, ww = h.shape[-2:]
        h = _to_map(_cross_attention(p.sub('cross_attn.fn'), _to_seq(h), cond, CROSS_HEADS), hh, ww) + h
    h = _block(p.sub('block2'), h, cfg.resnet_groups, scale_shift)
    if p.has('gca.to_k.weight'):
        h = h * _gca(p.sub('gca'), h)
    res = F.conv2d(x, p['res_conv.weight'], p['res_conv.bias']) if p.has('res_conv.weight') else x
    out = h + res
    if taps is not None:
        taps[name] = out
    return out


def _transformer(p: _P, x, context, cfg: UnetConfig):
    hh, ww = x.shape[-2:]
    x = _to_map(_attention(p.sub('layers.0.0.fn'), _to_seq(x), context, cfg.attn_heads), hh, ww) + x
    f = p.sub('layers.0.1')
    y = _ln(x, f['0.g'], dim=1)
    y = F.conv2d(y, f['1.weight'])
    y = F.gelu(y)
    y = _ln(y, f['3.g'], dim=1)
    y = F.conv2d(y, f['4.weight'])
    return y + x


def unet_forward(sd: Dict[str, torch.Tensor], cfg: UnetConfig, x: torch.Tensor, time: torch.Tensor,
                 cond_images: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """eps = Unet(x, log_snr, cond_images)   (imagen_pytorch.py:1470-1671, cond_drop_prob = 0).

    x [B,4,h,h]; time [B] (log-SNR, i.e. noise_scheduler.get_condition(t)); cond_images [B,256,h,h].
    ``taps`` (optional dict) receives named intermediate activations for layer-by-layer parity.
    """
    p = _P(sd)
    n = len(cfg.dim_mults)
    if cond_images.shape[-1] != x.shape[-1]:
        cond_images = F.interpolate(cond_images, x.shape[-1], mode='nearest')
    x = torch.cat((cond_images, x), dim=1)  # :1504 (keep mask is all ones at prob 0)
    ks = sorted(cfg.init_cross_embed_kernel_sizes)
    x = torch.cat([F.conv2d(x, p[f'init_conv.convs.{i}.weight'], p[f'init_conv.convs.{i}.bias'], padding=(k - 1) // 2)
                   for i, k in enumerate(ks)], dim=1)
    if taps is not None:
        taps['init_conv'] = x
    # time conditioning :1517-1522, :1600-1604
    tt = time[:, None]
    freqs = tt * p['to_time_hiddens.0.weights'][None, :] * 2 * math.pi
    four = torch.cat((tt, freqs.sin(), freqs.cos()), dim=-1)
    th = F.silu(F.linear(four, p['to_time_hiddens.1.weight'], p['to_time_hiddens.1.bias']))
    tokens = F.linear(th, p['to_time_tokens.0.weight'], p['to_time_tokens.0.bias']).view(x.shape[0], cfg.num_time_tokens, -1)
    t = F.linear(th, p['to_time_cond.0.weight'], p['to_time_cond.0.bias'])
    c = F.layer_norm(tokens, (tokens.shape[-1],), p['norm_cond.weight'], p['norm_cond.bias'])
    if taps is not None:
        taps['t'] = t
        taps['c'] = c

    hiddens: List[torch.Tensor] = []
    for i in range(n):
        last = i == n - 1
        x = _resnet(p.sub(f'downs.{i}.1'), x, t, c, cfg, taps, f'downs.{i}.1')
        for j in range(cfg.num_resnet_blocks[i]):
            x = _resnet(p.sub(f'downs.{i}.2.{j}'), x, t, None, cfg, taps, f'downs.{i}.2.{j}')
            hiddens.append(x)
        if cfg.layer_attns[i]:
            x = _transformer(p.sub(f'downs.{i}.3'), x, c, cfg)
            if taps is not None:
                taps[f'downs.{i}.3'] = x
        hiddens.append(x)
        if not last:
            x = F.conv2d(x, p[f'downs.{i}.4.weight'], p[f'downs.{i}.4.bias'], stride=2, padding=1)
        else:
            x = (F.conv2d(x, p[f'downs.{i}.4.fns.0.weight'], p[f'downs.{i}.4.fns.0.bias'], padding=1)
                 + F.conv2d(x, p[f'downs.{i}.4.fns.1.weight'], p[f'downs.{i}.4.fns.1.bias']))
        if taps is not None:
            taps[f'downs.{i}.4'] = x

    x = _resnet(p.sub('mid_block1'), x, t, c, cfg, taps, 'mid_block1')
    hh, ww = x.shape[-2:]
    x = _to_map(_attention(p.sub('mid_attn.fn.fn'), _to_seq(x), None, cfg.attn_heads), hh, ww) + x
    if taps is not None:
        taps['mid_attn'] = x
    x = _resnet(p.sub('mid_block2'), x, t, c, cfg, taps, 'mid_block2')

    skip_scale = 2 ** -0.5
    for i in range(n):
        last = i == n - 1
        ri = n - 1 - i
        x = torch.cat((x, hiddens.pop() * skip_scale), dim=1)
        x = _resnet(p.sub(f'ups.{i}.0'), x, t, c, cfg, taps, f'ups.{i}.0')
        for j in range(cfg.num_resnet_blocks[ri]):
            x = torch.cat((x, hiddens.pop() * skip_scale), dim=1)
            x = _resnet(p.sub(f'ups.{i}.1.{j}'), x, t, None, cfg, taps, f'ups.{i}.1.{j}')
        if cfg.layer_attns[ri]:
            x = _transformer(p.sub(f'ups.{i}.2'), x, c, cfg)
            if taps is not None:
                taps[f'ups.{i}.2'] = x
        if not last:
            x = F.pixel_shuffle(F.silu(F.conv2d(x, p[f'ups.{i}.3.net.0.weight'], p[f'ups.{i}.3.net.0.bias'])), 2)
            if taps is not None:
                taps[f'ups.{i}.3'] = x
    x = _resnet(p.sub('final_res_block'), x, t, None, cfg, taps, 'final_res_block')
    return F.conv2d(x, p['final_conv.weight'], p['final_conv.bias'], padding=1)


# ---------------------------------------------------------------------------------------------
# PLMS sampler (external/plms.py)
# ---------------------------------------------------------------------------------------------
class NoiseSource:
    """Stands in for every ``torch.randn_like`` the reference sampler calls, in call order
    (SURVEY.md Appendix C): one draw in ``plms_sample_loop`` then one per ``get_model_output``."""

    def __init__(self, seed: int = 0):
        self.seed, self.count = seed, 0

    def __call__(self, like: torch.Tensor) -> torch.Tensor:
        rng = np.random.default_rng([self.seed, self.count])
        self.count += 1
        return torch.from_numpy(rng.standard_normal(size=tuple(like.shape), dtype=np.float32)).to(like.dtype)


def plms_n_steps(max_thres: float, plms_steps: int = 50) -> int:
    return min(int(max_thres * plms_steps * 2), plms_steps)  # plms.py:87


def plms_sample(eps_fn: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], image: torch.Tensor,
                max_thres: float, noise: Callable[[torch.Tensor], torch.Tensor], plms_steps: int = 50,
                clip_value: float = 10.0):
    """Restates ``PLMSSampler.sample(image, max_thres, cond_images, return_noise=True)`` for
    max_thres < 0.99 (the only branch distillation.py:303 reaches: ``clamp(max=0.99)`` makes
    ``max_thres >= .99`` possible only at exactly 0.99, handled too).

    ``eps_fn(x, log_snr)`` is the UNet (cond_images bound by the caller).  Returns
    ``(img, x_noisy, noise0, alpha_cumprod, n_unet_calls)``.
    """
    b = image.shape[0]
    dt = image.dtype
    calls = 0
    if max_thres >= .99:  # plms.py:80-85
        lin = torch.linspace(1., 0., plms_steps + 1, dtype=torch.float32)
        n0 = noise(image)
        x_noisy, log_snr = q_sample(image, torch.full((b,), max_thres, dtype=dt), n0)
        img = image
    else:
        n_steps = plms_n_steps(max_thres, plms_steps)
        lin = torch.linspace(max_thres, 0.0, n_steps + 1, dtype=torch.float32)
        n0 = noise(image)
        img, log_snr = q_sample(image, torch.full((b,), max_thres, dtype=dt), n0)
        x_noisy = img
    pairs = [(lin[i].to(dt).expand(b), lin[i + 1].to(dt).expand(b)) for i in range(len(lin) - 1)]

    def model_output(x, t, t_next, pred_e=None):
        nonlocal calls
        if pred_e is None:
            pred_e = eps_fn(x, alpha_cosine_log_snr(t))
            calls += 1
        x_start = predict_start_from_noise(x, t, pred_e).clamp(-clip_value, clip_value)
        mean, _, logvar = q_posterior(x_start, x, t, t_next)
        nz = noise(x)
        mask = (1 - (t_next == 0).to(dt)).view(b, 1, 1, 1)
        return mean + mask * (0.5 * logvar).exp() * nz, x_start, pred_e

    old_eps: List[torch.Tensor] = []
    for t, t_next in pairs:
        _, _, e_t = model_output(img, t, t_next)
        if len(old_eps) == 0:
            x_prev, _, _ = model_output(img, t, t_next, pred_e=e_t)
            _, _, e_next = model_output(x_prev, t_next, t_next)
            e_prime = (e_t + e_next) / 2
        elif len(old_eps) == 1:
            e_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        img, _, _ = model_output(img, t, t_next, pred_e=e_prime)
        old_eps.append(e_t)
        if len(old_eps) >= 4:
            old_eps.pop(0)
    img = img.clamp(-clip_value, clip_value)
    return img, x_noisy, n0, torch.sigmoid(log_snr), calls
