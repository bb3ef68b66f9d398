"""CPU measurements behind two tolerances of the GPU parity tests (so they are checkable without a GPU):

* the bf16 discriminator input gradient is accepted at cosine > 0.99 / rel-L2 < 0.15 (tests/test_gpu_disc.py): bf16 rounding of
  operands and layer outputs ALONE -- the oracle itself with ``emulate_bf16`` -- moves that gradient by several per cent,
  while the logits move by ~1e-2;
* gradients through ReLU / max-pool stacks (the VGG16 of the perceptual term) are limited by GATE FLIPS: a forward
  perturbation of relative size delta changes the gradient by about sqrt(delta), not delta (tests/test_gpu_lpips.py header).
"""
import pytest
import torch


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize('cfg', ['tiny', 'small'])
def test_bf16_rounding_alone_moves_the_discriminator_gradient(cfg):
    from oracle import sg2_disc, synthetic
    wl = synthetic.make_workload(cfg, noise_strength=0.1)
    c = wl['cfg']
    D = sg2_disc.make_discriminator(img_resolution=c['img_resolution'], img_channels=c['img_channels'],
                                    channel_base=c['channel_base'], channel_max=c['channel_max'])
    B = wl['w0'].shape[0]
    x = (torch.rand([B, c['img_channels'], c['img_resolution'], c['img_resolution']], generator=torch.Generator().manual_seed(11)) * 2 - 1)
    x = x.requires_grad_(True)

    def run():
        x.grad = None
        logits = D(x, c=None)
        (torch.nn.functional.softplus(-logits).mean() * 0.7).backward()
        return logits.detach().clone(), x.grad.clone()
    l32, g32 = run()
    sg2_disc.Conv2dLayer.emulate_bf16 = True
    try:
        l16, g16 = run()
    finally:
        sg2_disc.Conv2dLayer.emulate_bf16 = False
    el = float((l16 - l32).abs().max() / l32.abs().max())
    eg = _rel(g16, g32)
    cos = float((g16.double() * g32.double()).sum() / g16.double().norm() / g32.double().norm())
    print(f'\n[{cfg}] bf16-emulating oracle vs fp32 oracle: logits max-rel {el:.3e}, input gradient rel-L2 {eg:.3e}, cosine {cos:.5f}')
    assert el < 2e-2                      # the forward error is the size one expects of bf16
    assert 5e-3 < eg < 0.15 and cos > 0.99   # ... the gradient error is several times larger: sign flips of near-zero pre-activations


def test_gate_flips_make_gradient_errors_scale_like_the_square_root_of_the_forward_error():
    from oracle import lpips as olp
    taps = olp.TAPS_INTREE
    st = {k: v.double() for k, v in olp.random_vgg_state(7, taps).items()}
    lw = olp.lin_weights(st, taps)
    g = torch.Generator().manual_seed(2)
    x0 = (torch.rand([2, 3, 64, 64], generator=g, dtype=torch.float64) * 2 - 1)
    bank = olp.vgg_features(st, torch.rand([3, 3, 64, 64], generator=g, dtype=torch.float64) * 2 - 1, taps)

    def grad(state):
        x = x0.clone().requires_grad_(True)
        olp.pair_distance(olp.vgg_features(state, x, taps), bank, lw).sum().backward()
        return x.grad
    g_ref = grad(st)
    errs = {}
    for delta in (1e-6, 1e-4):
        noisy = {k: v * (1.0 + delta * torch.randn(v.shape, generator=g, dtype=torch.float64)) if k.startswith('features') else v for k, v in st.items()}
        errs[delta] = _rel(grad(noisy), g_ref)
    print(f'\nweights perturbed by 1e-6 / 1e-4 (relative): gradient rel-L2 {errs[1e-6]:.3e} / {errs[1e-4]:.3e}')
    # 100 x the perturbation gives about 10 x the gradient error (sqrt law), far from 100 x; and the gradient error is orders
    # of magnitude above the perturbation itself
    ratio = errs[1e-4] / errs[1e-6]
    assert 3.0 < ratio < 40.0, errs
    assert errs[1e-6] > 20 * 1e-6, errs
