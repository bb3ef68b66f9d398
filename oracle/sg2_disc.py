"""Oracle restatement of the StyleGAN2 discriminator behind the reference's realism term
``calc_loss_disc`` (``augments/utils/util_latent_aug.py:363-371``: ``softplus(-D(x, c=None)).mean() * w_disc``).

TEST INFRASTRUCTURE (see oracle/__init__.py).

Like the generator, the class is NOT in /root/reference (unpickled third-party source,
NVlabs/stylegan3 ``training/networks_stylegan2.py``, unpinned).  This file restates the published
'resnet' architecture and is constrained by what IS in the tree:
 * constructor kwargs and defaults: ``models/stylegan3/legacy.py:220-250``
 * module tree and parameter names ``b{res}.{fromrgb,conv0,conv1,skip}.{weight,bias}``,
   ``b4.{conv,fc,out}.{weight,bias}``, ``resample_filter``: ``models/stylegan3/legacy.py:267-287``
 * the ops it composes: oracle/ops.py (``conv2d_resample`` incl. the down-sampling branches
   conv2d_resample.py:94-97,106-109, ``bias_act``), each pinned against the in-tree ref op
 * how the loop calls it: ``self.D(x, c=None)`` -> logits ``[B, 1]``.
Parity of the class composition itself is therefore UNPINNED by the reference.
"""
import math

import torch

from . import ops
from .sg2 import FC


class _RoundBf16(torch.autograd.Function):
    """Straight-through bf16 rounding of a layer output (and of the gradient flowing back through it)."""

    @staticmethod
    def forward(ctx, t):
        return t.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class Conv2dLayer(torch.nn.Module):
    """Equalised-lr conv + optional FIR down-sampling + bias + activation (names: weight, bias).

    ``emulate_bf16`` (class attribute, tests only): round the scaled weights and the layer output to bf16 --
    the arithmetic of the CUDA discriminator -- so that its gradient can be checked at a tolerance below the
    5 % that bf16 rounding alone moves the input gradient of a random-init D (sign flips of near-zero
    pre-activations)."""
    emulate_bf16 = False

    def __init__(self, cin, cout, k, bias=True, act='linear', down=1, conv_clamp=None, fir=(1, 3, 3, 1)):
        super().__init__()
        self.act, self.down, self.conv_clamp = act, down, conv_clamp
        self.padding = k // 2
        self.w_gain = 1.0 / math.sqrt(cin * k * k)
        self.weight = torch.nn.Parameter(torch.randn([cout, cin, k, k]))
        self.bias = torch.nn.Parameter(torch.zeros([cout])) if bias else None
        self.register_buffer('resample_filter', ops.setup_filter(list(fir)))

    def forward(self, x, gain=1.0):
        w = self.weight * self.w_gain
        if Conv2dLayer.emulate_bf16:
            w = w.bfloat16().float()
        x = ops.conv2d_resample(x, w, f=self.resample_filter, down=self.down, padding=self.padding, flip_weight=True)
        act_gain = ops._ACTS[self.act][2] * gain
        clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        x = ops.bias_act(x, self.bias, act=self.act, gain=act_gain, clamp=clamp)
        return _RoundBf16.apply(x) if Conv2dLayer.emulate_bf16 else x


class DiscBlock(torch.nn.Module):
    """'resnet' block: y = skip(x)*sqrt(.5) + conv1(conv0(x))*sqrt(.5); the top block owns ``fromrgb``."""

    def __init__(self, cin, ctmp, cout, res, img_channels, conv_clamp=256.0):
        super().__init__()
        self.cin, self.res = cin, res
        if cin == 0:
            self.fromrgb = Conv2dLayer(img_channels, ctmp, 1, act='lrelu', conv_clamp=conv_clamp)
        self.conv0 = Conv2dLayer(ctmp, ctmp, 3, act='lrelu', conv_clamp=conv_clamp)
        self.conv1 = Conv2dLayer(ctmp, cout, 3, act='lrelu', down=2, conv_clamp=conv_clamp)
        self.skip = Conv2dLayer(ctmp, cout, 1, bias=False, down=2)

    def forward(self, x, img):
        if self.cin == 0:
            x = self.fromrgb(img)
        y = self.skip(x, gain=math.sqrt(0.5))
        x = self.conv0(x)
        x = self.conv1(x, gain=math.sqrt(0.5))
        return y + x


class MinibatchStd(torch.nn.Module):
    def __init__(self, group_size=4, num_channels=1):
        super().__init__()
        self.group_size, self.num_channels = group_size, num_channels

    def forward(self, x):
        N, C, H, W = x.shape
        G = min(self.group_size, N) if self.group_size is not None else N
        F = self.num_channels
        c = C // F
        y = x.reshape(G, -1, F, c, H, W)
        y = y - y.mean(dim=0)
        y = y.square().mean(dim=0)
        y = (y + 1e-8).sqrt()
        y = y.mean(dim=[2, 3, 4])
        y = y.reshape(-1, F, 1, 1).repeat(G, 1, H, W)
        return torch.cat([x, y], dim=1)


class DiscEpilogue(torch.nn.Module):
    def __init__(self, cin, res=4, mbstd_group_size=4, mbstd_num_channels=1, conv_clamp=256.0):
        super().__init__()
        self.mbstd = MinibatchStd(mbstd_group_size, mbstd_num_channels)
        self.conv = Conv2dLayer(cin + mbstd_num_channels, cin, 3, act='lrelu', conv_clamp=conv_clamp)
        self.fc = FC(cin * res * res, cin, act='lrelu')
        self.out = FC(cin, 1)

    def forward(self, x):
        x = self.mbstd(x)
        x = self.conv(x)
        x = self.fc(x.flatten(1))
        return self.out(x)


class Discriminator(torch.nn.Module):
    def __init__(self, img_resolution, img_channels, channel_base=32768, channel_max=512, conv_clamp=256.0,
                 mbstd_group_size=4):
        super().__init__()
        self.img_resolution, self.img_channels = img_resolution, img_channels
        log2 = int(math.log2(img_resolution))
        self.block_resolutions = [2 ** i for i in range(log2, 2, -1)]
        ch = {r: min(channel_base // r, channel_max) for r in self.block_resolutions + [4]}
        self.channels = ch
        for r in self.block_resolutions:
            cin = ch[r] if r < img_resolution else 0
            setattr(self, f'b{r}', DiscBlock(cin, ch[r], ch[r // 2], r, img_channels, conv_clamp))
        self.b4 = DiscEpilogue(ch[4], mbstd_group_size=mbstd_group_size, conv_clamp=conv_clamp)

    def forward(self, img, c=None):
        x = None
        for r in self.block_resolutions:
            x = getattr(self, f'b{r}')(x, img)
        return self.b4(x)


def make_discriminator(seed=5, **kw):
    """Random-init D (weights ~ N(0,1), biases 0 as upstream; seeded)."""
    g = torch.Generator().manual_seed(seed)
    D = Discriminator(**kw)
    with torch.no_grad():
        for name, p in D.named_parameters():
            if name.endswith('weight'):
                p.copy_(torch.randn(p.shape, generator=g))
    return D.eval().requires_grad_(False)
