"""AutoencoderKL -- host-side mirror of the SD-v1 KL-f8 VAE the distillation loop calls every SDS step
(external/ldm/models/autoencoder.py:285-333, external/ldm/modules/diffusionmodules/model.py:82-142,150-203,368-568,
external/ldm/modules/distributions/distributions.py:24-62, config external/ldm/configs/sd-vae.yaml).

STATUS: this is SURVEY.md §8f "next" row #1, not yet ported to the sm_100a conv engine.  It is kept in plain torch
(cuDNN) so that a distillation step can run end to end with all of the reference's in-loop work present
(``vae.encode(...).mode()`` at sparsefusion/distillation.py:299 and ``vae.decode(...)`` at :309, both under no_grad);
DESIGN.md lists it as out of the B200-native scope of this round.  Same module tree and state_dict keys as the
reference (``encoder.*``, ``decoder.*``, ``quant_conv``, ``post_quant_conv``), without the pytorch_lightning / taming
base classes the reference drags in.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


def _swish(x):
    return x * torch.sigmoid(x)


def _norm(c):
    return nn.GroupNorm(num_groups=32, num_channels=c, eps=1e-6, affine=True)


class ResnetBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = _norm(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2 = _norm(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)

    def forward(self, x):
        h = self.conv1(_swish(self.norm1(x)))
        h = self.conv2(_swish(self.norm2(h)))
        if self.in_channels != self.out_channels:
            x = self.nin_shortcut(x)
        return x + h


class AttnBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.norm = _norm(c)
        self.q, self.k, self.v, self.proj_out = (nn.Conv2d(c, c, 1) for _ in range(4))

    def forward(self, x):
        h = self.norm(x)
        q, k, v = self.q(h), self.k(h), self.v(h)
        b, c, hh, ww = q.shape
        w_ = torch.bmm(q.reshape(b, c, hh * ww).permute(0, 2, 1), k.reshape(b, c, hh * ww)) * (int(c) ** (-0.5))
        w_ = F.softmax(w_, dim=2)
        h = torch.bmm(v.reshape(b, c, hh * ww), w_.permute(0, 2, 1)).reshape(b, c, hh, ww)
        return x + self.proj_out(h)


class Downsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 2, 0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1), mode='constant', value=0))  # asymmetric padding (model.py:73-75)


class Upsample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 1, 1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode='nearest'))


class _Level(nn.Module):
    pass


class Encoder(nn.Module):
    def __init__(self, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, z_channels=4, double_z=True):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i in range(self.num_resolutions):
            lvl = _Level()
            block_in, block_out = ch * in_ch_mult[i], ch * ch_mult[i]
            blocks = []
            for _ in range(num_res_blocks):
                blocks.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            lvl.block = nn.ModuleList(blocks)
            lvl.attn = nn.ModuleList()
            if i != self.num_resolutions - 1:
                lvl.downsample = Downsample(block_in)
            self.down.append(lvl)
        self.mid = _Level()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, 1, 1)

    def forward(self, x):
        h = self.conv_in(x)
        for i in range(self.num_resolutions):
            for blk in self.down[i].block:
                h = blk(h)
            if i != self.num_resolutions - 1:
                h = self.down[i].downsample(h)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        return self.conv_out(_swish(self.norm_out(h)))


class Decoder(nn.Module):
    def __init__(self, ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4):
        super().__init__()
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        block_in = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = _Level()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        ups = []
        for i in reversed(range(self.num_resolutions)):
            lvl = _Level()
            block_out = ch * ch_mult[i]
            blocks = []
            for _ in range(num_res_blocks + 1):
                blocks.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            lvl.block = nn.ModuleList(blocks)
            lvl.attn = nn.ModuleList()
            if i != 0:
                lvl.upsample = Upsample(block_in)
            ups.insert(0, lvl)
        self.up = nn.ModuleList(ups)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)

    def forward(self, z):
        h = self.conv_in(z)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        for i in reversed(range(self.num_resolutions)):
            for blk in self.up[i].block:
                h = blk(h)
            if i != 0:
                h = self.up[i].upsample(h)
        return self.conv_out(_swish(self.norm_out(h)))


class DiagonalGaussianDistribution:
    """distributions.py:24-62 (only what the loop uses: mode / sample)"""

    def __init__(self, parameters):
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def mode(self):
        return self.mean

    def sample(self):
        return self.mean + self.std * torch.randn_like(self.mean)


class AutoencoderKL(nn.Module):
    def __init__(self, ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, out_ch=3, z_channels=4, embed_dim=4):
        super().__init__()
        self.encoder = Encoder(ch, ch_mult, num_res_blocks, in_channels, z_channels, True)
        self.decoder = Decoder(ch, out_ch, ch_mult, num_res_blocks, z_channels)
        self.quant_conv = nn.Conv2d(2 * z_channels, 2 * embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, z_channels, 1)
        self.embed_dim = embed_dim

    def encode(self, x):
        return DiagonalGaussianDistribution(self.quant_conv(self.encoder(x)))

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z))

    def forward(self, x, sample_posterior=True):
        post = self.encode(x)
        z = post.sample() if sample_posterior else post.mode()
        return self.decode(z), post
