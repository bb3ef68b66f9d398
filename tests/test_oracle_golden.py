"""Pins the CPU oracle against outputs of the REFERENCE's own code (tests/golden/*.pt,
made by oracle/make_golden.py in the build container).  CPU only."""
import hashlib
import random

import pytest
import torch

from oracle import latent_aug as ola
from oracle import ops, synthetic
from conftest import rel_l2


def _digest(*ts):
    h = hashlib.sha256()
    for t in ts:
        h.update(t.detach().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def test_bias_act_matches_reference(golden):
    g = golden('ops.pt')['bias_act']
    for c in g['cases']:
        x = g['x'].clone().requires_grad_(True)
        y = ops.bias_act(x, g['b'], act=c['act'], gain=c['gain'], clamp=c['clamp'])
        dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5))
        (dx,) = torch.autograd.grad(y, x, dy)
        assert torch.equal(y.detach(), c['y']), c['act']
        assert torch.equal(dx, c['dx']), c['act']


def test_setup_filter_matches_reference(golden):
    g = golden('ops.pt')
    assert torch.equal(ops.setup_filter([1, 3, 3, 1]), g['setup_filter_1331'])
    assert torch.equal(ops.setup_filter(list(range(1, 13))), g['setup_filter_sep12'])


def test_upfirdn2d_matches_reference(golden):
    g = golden('ops.pt')['upfirdn2d']
    for c in g['cases']:
        x = g['x'].clone().requires_grad_(True)
        y = ops.upfirdn2d(x, g['f'], **c['kw'])
        dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(6))
        (dx,) = torch.autograd.grad(y, x, dy)
        assert y.shape == c['y'].shape
        torch.testing.assert_close(y.detach(), c['y'], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(dx, c['dx'], rtol=1e-6, atol=1e-6)
    u = golden('ops.pt')['upsample2d']
    torch.testing.assert_close(ops.upsample2d(u['x'], g['f']), u['y'], rtol=1e-6, atol=1e-6)


def test_conv2d_resample_matches_reference(golden):
    g = golden('ops.pt')['conv2d_resample']
    for c in g['cases']:
        kw = c['kw']
        x = g['x'].clone().requires_grad_(True)
        f = g['f'] if (kw.get('up', 1) > 1 or kw.get('down', 1) > 1) else None
        y = ops.conv2d_resample(x, c['w'], f=f, **kw)
        dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(7))
        (dx,) = torch.autograd.grad(y, x, dy)
        assert y.shape == c['y'].shape, kw
        torch.testing.assert_close(y.detach(), c['y'], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(dx, c['dx'], rtol=1e-5, atol=1e-5)


def test_fma_matches_reference(golden):
    g = golden('ops.pt')['fma']
    assert torch.equal(ops.fma(g['a'], g['b'], g['c']), g['y'])


def test_l2_loss_vectorized_matches_reference(golden):
    g = golden('losses.pt')
    for nd in (2, 3, 4):
        c = g[f'l2_{nd}d']
        D = ola.l2_loss_vectorized(c['X'], c['Y'], compute_mean=False)
        assert D.shape == c['D'].shape == (c['Y'].shape[0], c['X'].shape[0])   # [bank, batch]
        torch.testing.assert_close(D, c['D'], rtol=1e-6, atol=1e-5)
        torch.testing.assert_close(ola.l2_loss_vectorized(c['X'], c['Y']), c['mean'], rtol=1e-6, atol=1e-7)
    with pytest.raises(NotImplementedError):
        ola.l2_loss_vectorized(torch.zeros(3), torch.zeros(3))


def test_center_crop_bounds():
    assert ola.center_crop_bounds(256) == (38, 181)
    assert ola.center_crop_bounds(128) == (19, 90)
    assert ola.center_crop_bounds(512) == (75, 362)
    from torchvision import transforms
    x = torch.arange(32 * 32, dtype=torch.float32).reshape(1, 1, 32, 32)
    assert torch.equal(transforms.CenterCrop(ola.center_crop_bounds(32)[1])(x), ola.center_crop(x, 32))


@pytest.mark.parametrize('name', ['loop_tiny.pt', 'loop_tiny_soft.pt', 'loop_tiny128.pt', 'loop_small.pt'])
def test_loop_matches_reference(golden, name):
    g = golden(name)
    wl = synthetic.make_workload(g['config'], noise_strength=g['noise_strength'])
    assert _digest(wl['W'], wl['w0'], wl['X']) == g['inputs_digest'], 'synthetic inputs drifted'
    assert _digest(*wl['G'].state_dict().values()) == g['params_digest'], 'generator init drifted'
    G = wl['G']
    for fused in (True, False):
        orc = ola.LatentAugOracle(G, wl['W'], wl['X'], num_epochs=g['steps'], w_latent=g['w_latent'],
                                  w_pix=g['w_pix'], soft_aug=g['soft_aug'], alpha=g['alpha'], fused=fused)
        with torch.no_grad():
            ws0 = orc.broadcasting(wl['w0'])
            x0 = G.synthesis(ws0, noise_mode='const', fused=fused)
            assert rel_l2(x0, g['img0_const']) < 2e-5
            l_lat0 = float(ola.calc_loss_latent(ws0, wl['W'], g['w_latent']))
            l_pix0 = float(ola.calc_loss_pix(ola.center_crop(x0, orc.res), ola.center_crop(wl['X'], orc.res),
                                             g['w_pix'], orc.n_modalities))
            assert abs(l_lat0 - g['loss_latent0']) <= 1e-5 * abs(g['loss_latent0'])
            assert abs(l_pix0 - g['loss_pix0']) <= 1e-4 * abs(g['loss_pix0'])
            _, idx = ola.nearest_codes(ws0, wl['W'], k=g['nn_top4'].shape[1])
            assert torch.equal(idx[:, 0], g['nn_idx0'])
            assert torch.equal(idx, g['nn_top4'])
        random.seed(0)
        torch.manual_seed(1234)
        img, w_aug = orc.forward(wl['w0'].clone())
        # fused == the reference's formulation: tight.  Non-fused (the algebra the CUDA
        # path implements) differs by fp32 summation order, amplified by Adam's sign-like steps.
        tol_w, tol_img = (1e-5, 2e-4) if fused else (2e-4, 2e-3)
        assert rel_l2(w_aug[:, 0], g['w_aug']) < tol_w, (fused, rel_l2(w_aug[:, 0], g['w_aug']))
        assert rel_l2(img, g['img']) < tol_img, (fused, rel_l2(img, g['img']))
        assert torch.equal(w_aug[:, 0], w_aug[:, -1])


def test_lpips_oracle_matches_reference_golden(golden):
    """oracle/lpips.py against losses / image gradients / tap activations produced by the reference's own
    BaseNet + LinLayers + LPIPS.forward + crop pipeline (oracle/make_golden_lpips.py)."""
    from oracle import latent_aug as ola
    from oracle import lpips as olp
    G = golden('lpips.pt')
    for name, g in G.items():
        st = olp.random_vgg_state(7, g['taps'])
        img = g['img'].clone().requires_grad_(True)
        off, size = ola.center_crop_bounds(g['res'])
        xc = olp.crop(img[:, :, off:off + size, off:off + size], g['crop_pos'], g['crop_size'])
        bf = olp.bank_features(st, g['bank_crops'], g['taps'])
        loss = olp.calc_loss_lpips(st, xc, bf, g['w_lpips'], g['taps'], g['script'])
        (gr,) = torch.autograd.grad(loss, img)
        assert abs(float(loss) - g['loss']) < 1e-5 * abs(g['loss']), name
        assert rel_l2(gr, g['grad']) < 5e-4, name
        assert float(gr[:, :, :off].abs().max()) == 0.0          # nothing outside the centre crop
        for c in range(img.shape[1]):
            f = olp.vgg_features(st, xc[:, c:c + 1].repeat(1, 3, 1, 1).detach(), g['taps'])
            for k, t in enumerate(f):
                s, a = g['feat_sums'][c][k]
                assert abs(float(t.sum()) - s) < 1e-3 * a and abs(float(t.abs().sum()) - a) < 1e-4 * a


def test_four_term_loop_matches_reference_golden(golden):
    """The oracle loop with all four criteria against tests/golden/loop_four_terms.pt = the REFERENCE's own
    ``LatentAug.forward`` (crop draws, ``calc_loss_lpips_torchscript`` on LPIPS feature vectors, ``calc_loss_disc``, the sign
    combination, Adam) run by oracle/make_golden_four_terms.py; also pins the product's host-side bank-crop order."""
    from latentaugment_b200.augments.utils.util_latent_aug import feature_bank_crops
    from oracle import lpips as olp
    from oracle import sg2_disc
    g = golden('loop_four_terms.pt')
    cfg, wts = g['cfg'], g['weights']
    wl = synthetic.make_workload(dict(cfg), noise_strength=0.1)
    D = sg2_disc.make_discriminator(img_resolution=cfg['img_resolution'], img_channels=cfg['img_channels'],
                                    channel_base=cfg['channel_base'], channel_max=cfg['channel_max'])
    taps = tuple(g['taps'])
    st = olp.random_vgg_state(g['vgg_seed'], taps)
    random.seed(g['crop_seed'])
    crops = feature_bank_crops(wl['X'], cfg['img_resolution'], 64)
    assert torch.equal(crops, g['bank_crops']), 'bank windows are drawn in another order than the reference draws them'
    lp = dict(state=st, taps=taps, script=True, bank_feats=olp.bank_features(st, crops, taps))
    orc = ola.LatentAugOracle(wl['G'], wl['W'], wl['X'], num_epochs=cfg['steps'], D=D, lpips=lp, fused=True, **wts)
    random.seed(g['loop_seed'])
    torch.manual_seed(1234)
    img, w_aug = orc.forward(wl['w0'].clone())
    ref = g['losses']
    for t, (l_lat, l_pix, _, l_disc, l_lpips) in enumerate(orc.loss_log):
        for ours, theirs, tol in ((l_lat, ref[t, 0], 1e-5), (l_pix, ref[t, 1], 1e-4), (l_disc, ref[t, 2], 1e-4), (l_lpips, ref[t, 3], 1e-5)):
            assert abs(ours - float(theirs)) <= tol * abs(float(theirs)), (t, ours, float(theirs))
    assert rel_l2(w_aug[:, 0], g['w_aug']) < 1e-5, rel_l2(w_aug[:, 0], g['w_aug'])
    assert rel_l2(img, g['img']) < 2e-4, rel_l2(img, g['img'])
