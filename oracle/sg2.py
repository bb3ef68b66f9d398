"""Oracle restatement of the StyleGAN2 generator the reference loop drives.

TEST INFRASTRUCTURE (see oracle/__init__.py).

The network classes are NOT in /root/reference (they are unpickled third-party
source: NVlabs/stylegan3 ``training/networks_stylegan2.py``, unpinned).  This file
restates that published architecture and is constrained by what IS in the tree:
 * constructor kwargs and defaults: ``models/stylegan3/legacy.py:122-144``
 * module tree, parameter names, layouts and the ``affine.bias = 1`` init:
   ``models/stylegan3/legacy.py:171-203``
 * the ops it composes: oracle/ops.py (each pinned against the in-tree ref op)
 * how the loop calls it: ``G.z_dim / G.w_dim / G.num_ws``,
   ``G.mapping(z, c, truncation_psi=...)``, ``G.synthesis(ws, noise_mode=...)``
   (``augments/utils/util_latent_aug.py:119-121,203,227,460,488``).
"""
import math

import torch

from . import ops


class FC(torch.nn.Module):
    """Equalised-lr dense layer (legacy.py:172-176 names: ``weight``, ``bias``)."""

    def __init__(self, n_in, n_out, act='linear', lr_mul=1.0, bias_init=0.0):
        super().__init__()
        self.act = act
        self.weight = torch.nn.Parameter(torch.randn([n_out, n_in]) / lr_mul)
        self.bias = torch.nn.Parameter(torch.full([n_out], float(bias_init)))
        self.w_gain = lr_mul / math.sqrt(n_in)
        self.b_gain = lr_mul

    def forward(self, x):
        w = self.weight * self.w_gain
        b = self.bias * self.b_gain if self.b_gain != 1 else self.bias
        if self.act == 'linear':
            return torch.addmm(b.unsqueeze(0), x, w.t())
        return ops.bias_act(x.matmul(w.t()), b, act=self.act)


class Mapping(torch.nn.Module):
    """z -> ws.  SURVEY.md App. A.1; names ``fc{i}``, ``w_avg`` per legacy.py:172-176."""

    def __init__(self, z_dim, w_dim, num_ws, num_layers=8, lr_mul=0.01):
        super().__init__()
        self.z_dim, self.w_dim, self.num_ws, self.num_layers = z_dim, w_dim, num_ws, num_layers
        dims = [z_dim] + [w_dim] * num_layers
        for i in range(num_layers):
            setattr(self, f'fc{i}', FC(dims[i], dims[i + 1], act='lrelu', lr_mul=lr_mul))
        self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c=None, truncation_psi=1, truncation_cutoff=None):
        x = z.to(torch.float32)
        x = x * (x.square().mean(dim=1, keepdim=True) + 1e-8).rsqrt()
        for i in range(self.num_layers):
            x = getattr(self, f'fc{i}')(x)
        x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            if truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


def modulated_conv2d(x, weight, styles, noise=None, up=1, padding=0, resample_filter=None,
                     demodulate=True, flip_weight=True, fused=True):
    """SURVEY.md App. A.3.  ``fused=True`` is the eval-mode form the reference runs
    (per-sample weights, one grouped conv through conv2d_resample with groups=B);
    ``fused=False`` is the algebraically identical scale-conv-scale form the CUDA
    path implements (App. A.4)."""
    B = x.shape[0]
    O, I, kh, kw = weight.shape
    if fused:
        w = weight.unsqueeze(0) * styles.reshape(B, 1, I, 1, 1)
        if demodulate:
            d = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
            w = w * d.reshape(B, O, 1, 1, 1)
        y = ops.conv2d_resample(x.reshape(1, B * I, *x.shape[2:]), w.reshape(B * O, I, kh, kw),
                                f=resample_filter, up=up, padding=padding, groups=B, flip_weight=flip_weight)
        y = y.reshape(B, O, *y.shape[2:])
        if noise is not None:
            y = y + noise
        return y
    d = None
    if demodulate:
        w2 = weight.square().sum(dim=[2, 3])                       # [O,I]
        d = (styles.square().matmul(w2.t()) + 1e-8).rsqrt()        # [B,O]
    y = ops.conv2d_resample(x * styles.reshape(B, I, 1, 1), weight, f=resample_filter, up=up,
                            padding=padding, flip_weight=flip_weight)
    if d is not None and noise is not None:
        return ops.fma(y, d.reshape(B, O, 1, 1), noise)
    if d is not None:
        return y * d.reshape(B, O, 1, 1)
    if noise is not None:
        return y + noise
    return y


class SynthLayer(torch.nn.Module):
    """3x3 modulated conv (+ optional x2 up) + noise + bias + lrelu*sqrt2 + clamp.
    Names per legacy.py:178-195: weight, bias, noise_const, noise_strength, affine.*"""

    def __init__(self, cin, cout, w_dim, res, up=1, conv_clamp=256.0, fir=(1, 3, 3, 1)):
        super().__init__()
        self.res, self.up, self.conv_clamp = res, up, conv_clamp
        self.affine = FC(w_dim, cin, bias_init=1.0)
        self.weight = torch.nn.Parameter(torch.randn([cout, cin, 3, 3]))
        self.bias = torch.nn.Parameter(torch.zeros([cout]))
        self.register_buffer('noise_const', torch.randn([res, res]))
        self.noise_strength = torch.nn.Parameter(torch.zeros([]))
        self.register_buffer('resample_filter', ops.setup_filter(list(fir)))

    def forward(self, x, w, noise_mode='random', fused=True, gain=1.0):
        styles = self.affine(w)
        noise = None
        if noise_mode == 'random':
            noise = torch.randn([x.shape[0], 1, self.res, self.res], device=x.device) * self.noise_strength
        elif noise_mode == 'const':
            noise = self.noise_const * self.noise_strength
        x = modulated_conv2d(x, self.weight, styles, noise=noise, up=self.up, padding=1,
                             resample_filter=self.resample_filter, flip_weight=(self.up == 1), fused=fused)
        clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        return ops.bias_act(x, self.bias, act='lrelu', gain=ops.SQRT2 * gain, clamp=clamp)


class ToRGB(torch.nn.Module):
    """1x1 modulated conv without demodulation, linear, clamped (legacy.py:196-199)."""

    def __init__(self, cin, cout, w_dim, conv_clamp=256.0):
        super().__init__()
        self.conv_clamp = conv_clamp
        self.affine = FC(w_dim, cin, bias_init=1.0)
        self.weight = torch.nn.Parameter(torch.randn([cout, cin, 1, 1]))
        self.bias = torch.nn.Parameter(torch.zeros([cout]))
        self.w_gain = 1.0 / math.sqrt(cin)

    def forward(self, x, w, fused=True):
        styles = self.affine(w) * self.w_gain
        x = modulated_conv2d(x, self.weight, styles, demodulate=False, fused=fused)
        return ops.bias_act(x, self.bias, clamp=self.conv_clamp)


class SynthBlock(torch.nn.Module):
    def __init__(self, cin, cout, w_dim, res, img_channels, conv_clamp=256.0):
        super().__init__()
        self.cin, self.res = cin, res
        self.register_buffer('resample_filter', ops.setup_filter([1, 3, 3, 1]))
        self.num_conv = 0
        if cin == 0:
            self.const = torch.nn.Parameter(torch.randn([cout, res, res]))
        else:
            self.conv0 = SynthLayer(cin, cout, w_dim, res, up=2, conv_clamp=conv_clamp)
            self.num_conv += 1
        self.conv1 = SynthLayer(cout, cout, w_dim, res, conv_clamp=conv_clamp)
        self.num_conv += 1
        self.torgb = ToRGB(cout, img_channels, w_dim, conv_clamp=conv_clamp)
        self.num_torgb = 1

    def forward(self, x, img, ws, noise_mode='random', fused=True):
        wi = iter(ws.unbind(dim=1))
        if self.cin == 0:
            x = self.const.unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
        else:
            x = self.conv0(x, next(wi), noise_mode=noise_mode, fused=fused)
        x = self.conv1(x, next(wi), noise_mode=noise_mode, fused=fused)
        if img is not None:
            img = ops.upsample2d(img, self.resample_filter)
        y = self.torgb(x, next(wi), fused=fused).to(torch.float32)
        img = y if img is None else img + y
        return x, img


class Synthesis(torch.nn.Module):
    """Skip-architecture SG2 synthesis; blocks named ``b{res}``."""

    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, conv_clamp=256.0):
        super().__init__()
        self.w_dim, self.img_resolution, self.img_channels = w_dim, img_resolution, img_channels
        log2 = int(math.log2(img_resolution))
        assert 2 ** log2 == img_resolution and img_resolution >= 4
        self.block_resolutions = [2 ** i for i in range(2, log2 + 1)]
        ch = {r: min(channel_base // r, channel_max) for r in self.block_resolutions}
        self.channels = ch
        self.num_ws = 0
        for r in self.block_resolutions:
            blk = SynthBlock(ch[r // 2] if r > 4 else 0, ch[r], w_dim, r, img_channels, conv_clamp=conv_clamp)
            self.num_ws += blk.num_conv
            if r == img_resolution:
                self.num_ws += blk.num_torgb
            setattr(self, f'b{r}', blk)

    def forward(self, ws, noise_mode='random', fused=True):
        ws = ws.to(torch.float32)
        x = img = None
        idx = 0
        for r in self.block_resolutions:
            blk = getattr(self, f'b{r}')
            x, img = blk(x, img, ws.narrow(1, idx, blk.num_conv + blk.num_torgb), noise_mode=noise_mode, fused=fused)
            idx += blk.num_conv
        return img


class Generator(torch.nn.Module):
    def __init__(self, z_dim=512, w_dim=512, img_resolution=256, img_channels=3, channel_base=32768,
                 channel_max=512, conv_clamp=256.0, mapping_layers=8):
        super().__init__()
        self.z_dim, self.c_dim, self.w_dim = z_dim, 0, w_dim
        self.img_resolution, self.img_channels = img_resolution, img_channels
        self.synthesis = Synthesis(w_dim, img_resolution, img_channels, channel_base, channel_max, conv_clamp)
        self.num_ws = self.synthesis.num_ws
        self.mapping = Mapping(z_dim, w_dim, self.num_ws, num_layers=mapping_layers)

    def forward(self, z, c=None, truncation_psi=1, **kw):
        return self.synthesis(self.mapping(z, c, truncation_psi=truncation_psi), **kw)


def make_generator(seed=0, noise_strength=0.0, **kw):
    """Random-init generator per SURVEY.md §8(d): ``torch.manual_seed(seed)``, randn
    weights, zero biases, affine bias 1, w_avg 0; ``noise_strength`` set on every
    layer (0 for throughput runs, 0.1 for parity runs so the noise path is exercised)."""
    gen = torch.Generator().manual_seed(seed)
    state = torch.random.get_rng_state()
    torch.random.set_rng_state(gen.get_state())
    try:
        G = Generator(**kw)
    finally:
        torch.random.set_rng_state(state)
    with torch.no_grad():
        for name, p in G.named_parameters():
            if name.endswith('noise_strength'):
                p.fill_(noise_strength)
    return G.eval().requires_grad_(False)
