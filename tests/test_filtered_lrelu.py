"""filtered_lrelu (reference torch_utils/ops/filtered_lrelu.py) against golden vectors produced by the
REFERENCE's own ``_filtered_lrelu_ref`` and its autograd (tests/golden/filtered_lrelu.pt, oracle/make_golden_sg3.py).

CPU: the oracle restatement reproduces the golden outputs; the torch emulation of the CUDA kernel's call semantics,
driven with ``ops_sg3.backward_params``, reproduces the golden gradients (pins the adjoint-parameter algebra).
GPU: ``latentaugment_b200.ops_sg3.filtered_lrelu`` (through the C ABI) forward and backward, tolerance 1e-5 relative L2
(fp32 arithmetic; only the summation order differs)."""
import os

import pytest
import torch

from conftest import rel_l2

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'filtered_lrelu.pt')


def _cases():
    return torch.load(GOLDEN, weights_only=True)


def test_oracle_matches_reference_golden():
    from oracle import filtered_lrelu as ofl
    for g in _cases():
        c = g['cfg']
        y = ofl.filtered_lrelu_ref(g['x'], g['fu'], g['fd'], g['b'], c['up'], c['down'], c['padding'], 2 ** 0.5, 0.2, c['clamp'], c['flip'])
        assert y.shape == g['y'].shape
        assert rel_l2(y, g['y']) < 1e-6, c


def test_kernel_call_semantics_and_adjoint_parameters():
    """No GPU: emulate what la_filtered_lrelu computes, forward and (with the swapped parameters) backward."""
    from latentaugment_b200 import ops_sg3
    from oracle import filtered_lrelu as ofl
    for g in _cases():
        c = g['cfg']
        pad4 = ops_sg3._pad4(c['padding'])
        fu, fd = ops_sg3._taps(g['fu'], c['flip']), ops_sg3._taps(g['fd'], c['flip'])
        y, mask = ofl.emulate_kernel_call(g['x'].double(), fu, fd, None if g['b'] is None else g['b'].double(), c['up'], c['down'], pad4,
                                          2 ** 0.5, 0.2, c['clamp'], want_mask=True)
        assert rel_l2(y, g['y']) < 1e-6, c
        H, W = g['x'].shape[2:]
        bp = ops_sg3.backward_params(H, W, fu, fd, c['up'], c['down'], *pad4)
        gx = ofl.emulate_kernel_call(g['gy'].double(), bp['fu'], bp['fd'], None, bp['up'], bp['down'], bp['padding'], 2 ** 0.5, 0.2, None,
                                     mask_in=mask, mask_geom=(bp['mask_oy'], bp['mask_ox'], bp['mask_h'], bp['mask_w']))
        assert gx.shape == g['gx'].shape, c
        assert rel_l2(gx, g['gx']) < 1e-6, c
        if g['gb'] is not None:
            assert rel_l2(gx.sum(dim=[0, 2, 3]), g['gb']) < 1e-6


def test_c_abi_declares_the_operator():
    hdr = open(os.path.join(os.path.dirname(__file__), '..', 'include', 'latentaugment_b200.h')).read()
    assert 'int la_filtered_lrelu(' in hdr


@pytest.mark.gpu
def test_filtered_lrelu_gpu_matches_reference_golden():
    from latentaugment_b200.ops_sg3 import filtered_lrelu
    for g in _cases():
        c = g['cfg']
        x = g['x'].cuda().requires_grad_(True)
        b = g['b'].cuda().requires_grad_(True) if g['b'] is not None else None
        y = filtered_lrelu(x, g['fu'], g['fd'], b, up=c['up'], down=c['down'], padding=c['padding'], clamp=c['clamp'], flip_filter=c['flip'])
        (y * g['gy'].cuda()).sum().backward()
        ey, ex = rel_l2(y.detach().cpu(), g['y']), rel_l2(x.grad.cpu(), g['gx'])
        eb = rel_l2(b.grad.cpu(), g['gb']) if b is not None else 0.0
        print(f'\n[filtered_lrelu {c}] y={ey:.2e} gx={ex:.2e} gb={eb:.2e}')
        assert y.shape == g['y'].shape
        assert ey < 1e-5 and ex < 1e-5 and eb < 1e-5


@pytest.mark.gpu
def test_filtered_lrelu_gpu_sg3_layer_shape_vs_oracle():
    """SG3-T layer shape (12-tap up 2 / down 2 filters, 10-pixel margins) at 64 channels x 148^2: oracle parity + bandwidth."""
    from latentaugment_b200.ops_sg3 import filtered_lrelu
    from oracle import filtered_lrelu as ofl
    from oracle import ops
    gen = torch.Generator().manual_seed(0)
    x = torch.randn([2, 64, 148, 148], generator=gen)
    b = torch.randn([64], generator=gen)
    fu = ops.setup_filter((torch.rand(12, generator=gen) + 0.1).tolist())
    fd = ops.setup_filter((torch.rand(12, generator=gen) + 0.1).tolist())
    y_ref = ofl.filtered_lrelu_ref(x, fu, fd, b, 2, 2, 10, 2 ** 0.5, 0.2, 256.0)
    y = filtered_lrelu(x.cuda(), fu, fd, b.cuda(), up=2, down=2, padding=10, clamp=256.0)
    assert rel_l2(y.cpu(), y_ref) < 1e-5
    xc, bc = x.cuda().repeat(8, 1, 1, 1), b.cuda()
    for _ in range(3):
        filtered_lrelu(xc, fu, fd, bc, up=2, down=2, padding=10, clamp=256.0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        yc = filtered_lrelu(xc, fu, fd, bc, up=2, down=2, padding=10, clamp=256.0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = (xc.numel() + yc.numel()) * 4 / 1e9
    print(f'\n[filtered_lrelu 16x64x148^2] {ms:.3f} ms, {gb / (ms * 1e-3):.0f} GB/s of x + y traffic (+ int8 mask {yc.numel() * 4 / 1e9:.2f} GB)')
