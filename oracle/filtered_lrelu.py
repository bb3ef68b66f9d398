"""Oracle restatement of ``filtered_lrelu`` (reference ``torch_utils/ops/filtered_lrelu.py:121-153``, the ref path:
``bias_act -> upfirdn2d(up, gain=up^2) -> bias_act(lrelu, gain, clamp) -> upfirdn2d(down)``) and a torch emulation of the
CUDA kernel's call semantics used to check the adjoint-parameter algebra on the CPU.  TEST INFRASTRUCTURE."""
import torch
import torch.nn.functional as F

from . import ops


def filtered_lrelu_ref(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=2 ** 0.5, slope=0.2, clamp=None, flip_filter=False):
    px0, px1, py0, py1 = ops._pad4(padding)
    x = ops.bias_act(x, b)
    x = ops.upfirdn2d(x, fu, up=up, padding=[px0, px1, py0, py1], gain=up ** 2, flip_filter=flip_filter)
    x = ops.bias_act(x, act='lrelu', alpha=slope, gain=gain, clamp=clamp)
    return ops.upfirdn2d(x, fd, down=down, flip_filter=flip_filter)


def _corr_sep(z, taps):
    k = torch.tensor(taps, dtype=z.dtype)
    c = z.shape[1]
    z = F.conv2d(z, k[None, None, None, :].repeat(c, 1, 1, 1), groups=c)
    return F.conv2d(z, k[None, None, :, None].repeat(c, 1, 1, 1), groups=c)


def emulate_kernel_call(x, fu, fd, b, up, down, padding, gain, slope, clamp, mask_in=None, mask_geom=None, want_mask=False):
    """What ``la_filtered_lrelu`` computes (flip_filter=0), in torch: fu / fd are python lists of taps or None."""
    px0, px1, py0, py1 = padding
    fu = fu or [1.0]
    fd = fd or [1.0]
    n, c, h, w = x.shape
    if b is not None:
        x = x + b.reshape(1, -1, 1, 1)
    z = x.new_zeros([n, c, h, up, w, up])
    z[:, :, :, 0, :, 0] = x
    z = z.reshape(n, c, h * up, w * up)
    z = F.pad(z, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    z = z[:, :, max(-py0, 0): z.shape[2] - max(-py1, 0), max(-px0, 0): z.shape[3] - max(-px1, 0)]
    u = _corr_sep(z, [v * up for v in reversed(fu)])
    mask = None
    if mask_in is not None:
        oy, ox, mh, mw = mask_geom
        cls = torch.zeros(u.shape, dtype=torch.int8)
        ys, xs = max(oy, 0), max(ox, 0)
        ye, xe = min(oy + mh, u.shape[2]), min(ox + mw, u.shape[3])
        cls[:, :, ys:ye, xs:xe] = mask_in.reshape(n, c, mh, mw)[:, :, ys - oy:ye - oy, xs - ox:xe - ox]
        a = u * gain * torch.where(cls == 1, 1.0, torch.where(cls == 2, slope, 0.0))
    else:
        a = torch.where(u > 0, u, u * slope) * gain
        mask = torch.where(u > 0, 1, 2).to(torch.int8)
        if clamp is not None and clamp >= 0:
            sat = ~(a.abs() < clamp)
            a = a.clamp(-clamp, clamp)
            mask = torch.where(sat, torch.zeros_like(mask), mask)
    d = _corr_sep(a, list(reversed(fd)))
    y = d[:, :, ::down, ::down]
    return (y, mask.reshape(n * c, *mask.shape[2:])) if want_mask else y
