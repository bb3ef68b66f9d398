"""Oracle restatement of the reference's native-op layer (CPU, plain torch).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference files are under
``/root/reference/models/stylegan3/torch_utils/ops/``.
"""
import math

import torch
import torch.nn.functional as F

SQRT2 = math.sqrt(2.0)

# (function, default alpha, default gain) -- bias_act.py:21-31
_ACTS = {
    'linear': (lambda x, a: x, 0.0, 1.0),
    'relu': (lambda x, a: F.relu(x), 0.0, SQRT2),
    'lrelu': (lambda x, a: F.leaky_relu(x, a), 0.2, SQRT2),
    'tanh': (lambda x, a: torch.tanh(x), 0.0, 1.0),
    'sigmoid': (lambda x, a: torch.sigmoid(x), 0.0, 1.0),
    'elu': (lambda x, a: F.elu(x), 0.0, 1.0),
    'selu': (lambda x, a: F.selu(x), 0.0, 1.0),
    'softplus': (lambda x, a: F.softplus(x), 0.0, 1.0),
    'swish': (lambda x, a: torch.sigmoid(x) * x, 0.0, SQRT2),
}


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """``+b -> act -> *gain -> clamp``; follows ``_bias_act_ref`` (bias_act.py:91-120)."""
    fn, d_alpha, d_gain = _ACTS[act]
    alpha = d_alpha if alpha is None else float(alpha)
    gain = d_gain if gain is None else float(gain)
    if b is not None:
        shape = [1] * x.ndim
        shape[dim] = -1
        x = x + b.reshape(shape)
    x = fn(x, alpha)
    if gain != 1:
        x = x * gain
    if clamp is not None and clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def setup_filter(taps, normalize=True, flip_filter=False, gain=1.0, separable=None):
    """FIR preparation; follows ``setup_filter`` (upfirdn2d.py:70-114)."""
    f = torch.as_tensor(1.0 if taps is None else taps, dtype=torch.float32)
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = f.ndim == 1 and f.numel() >= 8
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    return f * (gain ** (f.ndim / 2))


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _pad4(p):
    if isinstance(p, int):
        return p, p, p, p
    p = list(p)
    if len(p) == 2:
        return p[0], p[0], p[1], p[1]
    return tuple(p)


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1.0):
    """zero-insert -> pad/crop -> FIR -> decimate; follows ``_upfirdn2d_ref``
    (upfirdn2d.py:167-211).  ``f`` None = identity."""
    n, c, h, w = x.shape
    ux, uy = _pair(up)
    dx, dy = _pair(down)
    px0, px1, py0, py1 = _pad4(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    # zero insertion (:187-189)
    z = x.new_zeros([n, c, h, uy, w, ux])
    z[:, :, :, 0, :, 0] = x
    z = z.reshape(n, c, h * uy, w * ux)
    # pad, then crop for negative paddings (:192-193)
    z = F.pad(z, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    z = z[:, :, max(-py0, 0): z.shape[2] - max(-py1, 0), max(-px0, 0): z.shape[3] - max(-px1, 0)]
    # filter: gain, dtype, true convolution unless flip_filter (:196-199)
    k = (f * (gain ** (f.ndim / 2))).to(z.dtype)
    if not flip_filter:
        k = k.flip(list(range(k.ndim)))
    if k.ndim == 2:
        z = F.conv2d(z, k[None, None].repeat(c, 1, 1, 1), groups=c)
    else:
        z = F.conv2d(z, k[None, None, None, :].repeat(c, 1, 1, 1), groups=c)
        z = F.conv2d(z, k[None, None, :, None].repeat(c, 1, 1, 1), groups=c)
    return z[:, :, ::dy, ::dx]


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1.0):
    """follows ``upsample2d`` (upfirdn2d.py:313-348): pad [(fw+up-1)//2, (fw-up)//2], gain*up^2."""
    ux, uy = _pair(up)
    px0, px1, py0, py1 = _pad4(padding)
    fh, fw = (f.shape[0], f.shape[-1]) if f.ndim == 2 else (f.shape[0], f.shape[0])
    p = [px0 + (fw + ux - 1) // 2, px1 + (fw - ux) // 2, py0 + (fh + uy - 1) // 2, py1 + (fh - uy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * ux * uy)


def _conv(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """follows ``_conv2d_wrapper`` (conv2d_resample.py:29-41): torch conv2d is a
    correlation, so flip_weight=False means "flip the taps first"."""
    if not flip_weight and (w.shape[2] > 1 or w.shape[3] > 1):
        w = w.flip([2, 3])
    op = F.conv_transpose2d if transpose else F.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """follows ``conv2d_resample`` (conv2d_resample.py:46-141), all branches."""
    oc, icg, kh, kw = w.shape
    if f is None:
        fw = fh = 1
    else:
        fw, fh = f.shape[-1], f.shape[0]
    px0, px1, py0, py1 = _pad4(padding)
    if up > 1:   # :82-86
        px0 += (fw + up - 1) // 2; px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2; py1 += (fh - up) // 2
    if down > 1:  # :87-91
        px0 += (fw - down + 1) // 2; px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2; py1 += (fh - down) // 2
    if kw == 1 and kh == 1 and down > 1 and up == 1:   # :94-97
        x = upfirdn2d(x, f, down=down, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, groups=groups, flip_weight=flip_weight)
    if kw == 1 and kh == 1 and up > 1 and down == 1:   # :100-103
        x = _conv(x, w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d(x, f, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    if down > 1 and up == 1:                            # :106-109
        x = upfirdn2d(x, f, padding=[px0, px1, py0, py1], flip_filter=flip_filter)
        return _conv(x, w, stride=down, groups=groups, flip_weight=flip_weight)
    if up > 1:                                          # :112-129
        if groups == 1:
            w = w.transpose(0, 1)
        else:
            w = w.reshape(groups, oc // groups, icg, kh, kw).transpose(1, 2)
            w = w.reshape(groups * icg, oc // groups, kh, kw)
        px0 -= kw - 1; px1 -= kw - up; py0 -= kh - 1; py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        x = _conv(x, w, stride=up, padding=[pyt, pxt], groups=groups, transpose=True, flip_weight=(not flip_weight))
        x = upfirdn2d(x, f, padding=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
        return x
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:   # :132-134
        return _conv(x, w, padding=[py0, px0], groups=groups, flip_weight=flip_weight)
    x = upfirdn2d(x, (f if up > 1 else None), up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)  # :137-141
    x = _conv(x, w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
    return x


def fma(a, b, c):
    """``a*b+c``; follows ``fma`` (fma.py:15-23)."""
    return torch.addcmul(c, a, b)
