"""Precision study (CPU, oracle only): how accurate must the tensor-core operands be for
the N-step loop to stay within rel-L2 1e-3 of the fp32 reference?

Emulates operand rounding of the conv GEMMs (forward A/B operands and the backward
data-gradient operands) inside the oracle's non-fused generator and reports the
relative L2 error of the final w and image against the exact-fp32 oracle run.

    python -m tools.precision_study [tiny small]
"""
import random
import sys

import torch

from oracle import latent_aug as ola
from oracle import ops, sg2, synthetic


def rnd(x, mode):
    if mode == 'fp32':
        return x
    if mode == 'bf16':
        return x.to(torch.bfloat16).to(torch.float32)
    if mode == 'bf16x2':          # hi + lo split: ~16 mantissa bits
        hi = x.to(torch.bfloat16).to(torch.float32)
        lo = (x - hi).to(torch.bfloat16).to(torch.float32)
        return hi + lo
    if mode == 'tf32':
        i = x.view(torch.int32)
        i = (i + 0x1000) & ~0x1FFF
        return i.view(torch.float32)
    raise ValueError(mode)


class Q(torch.autograd.Function):
    """round in forward with `fmode`, round the incoming gradient with `bmode`."""

    @staticmethod
    def forward(ctx, x, fmode, bmode):
        ctx.bmode = bmode
        return rnd(x, fmode)

    @staticmethod
    def backward(ctx, g):
        return rnd(g.contiguous(), ctx.bmode), None, None


def patched_modconv(mode, store):
    def modulated_conv2d(x, weight, styles, noise=None, up=1, padding=0, resample_filter=None,
                         demodulate=True, flip_weight=True, fused=True):
        B = x.shape[0]
        O, I, kh, kw = weight.shape
        d = None
        if demodulate:
            w2 = weight.square().sum(dim=[2, 3])
            d = (styles.square().matmul(w2.t()) + 1e-8).rsqrt()
        is_conv = kh > 1
        m = mode if is_conv else 'fp32'      # toRGB runs on fp32 SIMT
        xs = Q.apply(x * styles.reshape(B, I, 1, 1), m, 'fp32')
        y = ops.conv2d_resample(xs, rnd(weight, m), f=resample_filter, up=up, padding=padding, flip_weight=flip_weight)
        y = Q.apply(y, 'fp32', m)            # g_y operand of the dgrad GEMM
        if d is not None:
            y = y * d.reshape(B, O, 1, 1)
        if noise is not None:
            y = y + noise
        return y
    return modulated_conv2d


def run(cfg, mode, store, steps=None):
    g = synthetic.make_workload(cfg, noise_strength=0.1)
    c = g['cfg']
    orig = sg2.modulated_conv2d
    orig_ba = ops.bias_act
    sg2.modulated_conv2d = patched_modconv(mode, store)

    def ba(x, b=None, **kw):
        y = orig_ba(x, b, **kw)
        if store != 'fp32' and x.ndim == 4 and kw.get('act') == 'lrelu':
            y = Q.apply(y, store, 'fp32')
        return y
    ops.bias_act = ba
    try:
        orc = ola.LatentAugOracle(g['G'], g['W'], g['X'], num_epochs=steps or c['steps'], fused=False)
        random.seed(0)
        torch.manual_seed(1234)
        img, w = orc.forward(g['w0'].clone())
    finally:
        sg2.modulated_conv2d = orig
        ops.bias_act = orig_ba
    return img, w[:, 0], g['w0'][:, 0]


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


if __name__ == '__main__':
    cfgs = sys.argv[1:] or ['tiny', 'small']
    for cfg in cfgs:
        for steps in (None, 10):
            img0, w0, wi = run(cfg, 'fp32', 'fp32', steps)
            print(f'{cfg} steps={steps}: |w_aug - w_init| / |w_init| = {rel(w0, wi):.3e}')
            for mode, store in [('bf16', 'bf16'), ('bf16', 'fp32'), ('tf32', 'fp32'), ('bf16x2', 'bf16x2'), ('bf16x2', 'fp32')]:
                img, w, _ = run(cfg, mode, store, steps)
                print(f'  gemm={mode:7s} store={store:7s} rel_w={rel(w, w0):.3e} rel_img={rel(img, img0):.3e}')
