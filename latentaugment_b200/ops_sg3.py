"""StyleGAN3 operator ``filtered_lrelu`` on the GPU (reference
``models/stylegan3/torch_utils/ops/filtered_lrelu.py:56-153``).

``filtered_lrelu(x, fu, fd, b, up, down, padding, gain, slope, clamp, flip_filter)`` has the reference's signature and
semantics for separable (1-D) filters and is differentiable w.r.t. ``x`` and ``b``: the forward call records the
activation-derivative class of every intermediate pixel and the backward pass is the SAME kernel with the two filters'
roles swapped (``backward_params``).  There is no CPU path: tensors must live on an sm_100 device.
"""
import ctypes as C

import torch

from . import _lib


def _pad4(padding):
    if isinstance(padding, int):
        return padding, padding, padding, padding
    p = list(padding)
    if len(p) == 2:
        return p[0], p[0], p[1], p[1]
    return tuple(p)


def _taps(f, flip_filter):
    """1-D taps in the flip_filter=False convention (the op convolves), or None for the identity."""
    if f is None:
        return None
    f = torch.as_tensor(f, dtype=torch.float32).detach().cpu()
    if f.ndim == 2:      # setup_filter() turns short 1-D filters into outer(f, f): accept rank-1 symmetric 2-D filters
        f1 = f.sum(1) / f.sum().abs().sqrt().clamp_min(1e-30) * torch.sign(f.sum())
        if f.shape[0] != f.shape[1] or not torch.allclose(torch.outer(f1, f1), f, rtol=1e-5, atol=1e-7):
            raise NotImplementedError('filtered_lrelu: only separable filters (1-D, or outer(f, f)) are supported')
        f = f1
    if f.ndim != 1:
        raise NotImplementedError('filtered_lrelu: only separable filters are supported')
    return [float(v) for v in (f.flip(0) if flip_filter else f)]


def sizes(H, up, down, p0, p1, fu_taps, fd_taps):
    mid = H * up + p0 + p1 - (fu_taps - 1)
    out = (mid - (fd_taps - 1) + down - 1) // down
    return mid, out


def backward_params(H, W, fu, fd, up, down, px0, px1, py0, py1):
    """Arguments of the kernel call that maps dL/dy to dL/dx (per axis: input length L, pads p0/p1).

    Forward per axis:  U[m] = sum_k fuc[k] xup[m + k - p0],  y[o] = sum_k fdc[k] A[o*down + k]  (fuc / fdc = flipped taps).
    Adjoint: zero-insert g by `down`, correlate with the UNflipped fd (taps reversed and divided by the kernel's built-in
    up-gain), multiply by the recorded activation derivative, correlate with the unflipped fu * up and keep every
    `up`-th sample.  The intermediate of the adjoint call is the forward intermediate shifted by t = (Fu - 1) - p0."""
    Fu, Fd = len(fu) if fu else 1, len(fd) if fd else 1
    fu = fu or [1.0]
    fd = fd or [1.0]
    out = {}

    def axis(L, p0, p1):
        mid, o = sizes(L, up, down, p0, p1, Fu, Fd)
        t = (Fu - 1) - p0
        q0 = (Fd - 1) + t
        mid_b = (L - 1) * up + Fu
        q1 = mid_b + (Fd - 1) - o * down - q0
        return q0, q1, t, mid
    qx0, qx1, tx, mid_w = axis(W, px0, px1)
    qy0, qy1, ty, mid_h = axis(H, py0, py1)
    out.update(up=down, down=up, fu=[v / down for v in reversed(fd)], fd=[v * up for v in reversed(fu)],
               padding=(qx0, qx1, qy0, qy1), mask_oy=ty, mask_ox=tx, mask_h=mid_h, mask_w=mid_w)
    return out


def _call(x, fu, fd, b, up, down, padding, gain, slope, clamp, mask_in=None, mask_out=None, mask_geom=(0, 0, 0, 0)):
    lib = _lib.load()
    N, Cc, H, W = x.shape
    px0, px1, py0, py1 = padding
    Fu, Fd = len(fu) if fu else 1, len(fd) if fd else 1
    _, out_w = sizes(W, up, down, px0, px1, Fu, Fd)
    _, out_h = sizes(H, up, down, py0, py1, Fu, Fd)
    y = torch.empty([N, Cc, out_h, out_w], device=x.device, dtype=torch.float32)
    fu_a = (C.c_float * Fu)(*fu) if fu else None
    fd_a = (C.c_float * Fd)(*fd) if fd else None
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
    with torch.cuda.device(x.device):
        _lib.check(lib.la_filtered_lrelu(ptr(x), N, Cc, H, W, fu_a, Fu if fu else 0, fd_a, Fd if fd else 0, ptr(b), up, down,
                                         px0, px1, py0, py1, float(gain), float(slope), -1.0 if clamp is None else float(clamp), 0,
                                         ptr(mask_in), ptr(mask_out), *mask_geom, ptr(y),
                                         C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
    return y


class _FilteredLrelu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, b, fu, fd, up, down, padding, gain, slope, clamp):
        x = x.detach().float().contiguous()
        bb = b.detach().float().contiguous() if b is not None else None
        N, Cc, H, W = x.shape
        Fu = len(fu) if fu else 1
        mid_w = W * up + padding[0] + padding[1] - (Fu - 1)
        mid_h = H * up + padding[2] + padding[3] - (Fu - 1)
        mask = torch.empty([N * Cc, mid_h, mid_w], dtype=torch.int8, device=x.device)
        y = _call(x, fu, fd, bb, up, down, padding, gain, slope, clamp, mask_out=mask)
        ctx.save_for_backward(mask)
        ctx.meta = (x.shape, fu, fd, up, down, padding, gain, slope, b is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        (mask,) = ctx.saved_tensors
        shape, fu, fd, up, down, padding, gain, slope, has_b = ctx.meta
        bp = backward_params(shape[2], shape[3], fu, fd, up, down, *padding)
        gx = _call(gy.detach().float().contiguous(), bp['fu'], bp['fd'], None, bp['up'], bp['down'], bp['padding'], gain, slope, None,
                   mask_in=mask, mask_geom=(bp['mask_oy'], bp['mask_ox'], bp['mask_h'], bp['mask_w']))
        assert gx.shape == torch.Size(shape), (gx.shape, shape)
        gb = gx.sum(dim=[0, 2, 3]) if has_b else None
        return gx, gb, None, None, None, None, None, None, None, None


def filtered_lrelu(x, fu=None, fd=None, b=None, up=1, down=1, padding=0, gain=2 ** 0.5, slope=0.2, clamp=None, flip_filter=False):
    """Reference signature (filtered_lrelu.py:56).  x [N, C, H, W] fp32 on an sm_100 device."""
    if not x.is_cuda:
        raise _lib.LatentAugmentError('filtered_lrelu: latentaugment_b200 has no CPU path')
    return _FilteredLrelu.apply(x, b, _taps(fu, flip_filter), _taps(fd, flip_filter), int(up), int(down), _pad4(padding), gain, slope, clamp)
