"""GPU parity at the BENCHMARKED shapes: config C2 (StyleGAN2 256x256 3-ch, full channel table,
4096-code bank, 10 Adam steps) and config C3's generator (512x512), in both precisions.

Two checkers:
  * the reference's own ``LatentAug.forward`` outputs committed as ``tests/golden/loop_c2.pt`` (batch 8, 10 steps) and
    ``loop_c3.pt`` (batch 2, 2 steps) -- made by ``python -m oracle.make_golden --only loop_c2,loop_c3``;
  * the CPU oracle at C2's full batch of 32 x 10 steps, run on the box's host cores (about a minute).

Tolerances (relative L2): fp32_parity 1e-3 (north_star), bf16 1e-2 (the widened tolerance the north_star allows).
Besides the norms every test reports the Adam SIGN-FLIP count: at step 1 Adam moves each coordinate by lr * g/|g|,
so a gradient component that is zero to rounding lands a full 2*lr away from the checker's.  Such
components (|dw| > lr) are counted and the error of the remaining ones is asserted separately, so the norm gate is
not a coin toss on near-zero gradients.
"""
import json
import os
import random

import pytest
import torch

from conftest import ROOT, rel_l2

pytestmark = pytest.mark.gpu

TOL = {'fp32_parity': 1e-3, 'bf16': 1e-2}
LR = 0.01
_report = {}


def _engine(wl, precision, batch):
    from latentaugment_b200.engine import SynthesisEngine
    G = wl['G']
    return SynthesisEngine(dict(G.state_dict()), img_resolution=G.img_resolution, img_channels=G.img_channels,
                           w_dim=G.w_dim, z_dim=G.z_dim, batch=batch, precision=precision)


def flip_stats(w, w_ref, lr=LR):
    """(number of components further than lr from the checker, rel-L2 over the others, max |dw| over the others)."""
    d = (w.double() - w_ref.double()).abs()
    flips = d > lr
    keep = ~flips
    rest = float((d[keep].square().sum() / w_ref.double()[keep].square().sum().clamp_min(1e-30)).sqrt())
    return int(flips.sum()), rest, float(d[keep].max()) if keep.any() else 0.0


def _record(key, **kw):
    _report[key] = kw
    out = os.path.join(ROOT, 'gpurun_out')
    try:
        os.makedirs(out, exist_ok=True)
        json.dump(_report, open(os.path.join(out, 'parity_bench_shapes.json'), 'w'), indent=1)
    except OSError:
        pass


def _run(eng, wl, steps, noise=None):
    img, w_aug, losses = eng.augment(wl['w0'], num_steps=steps, lr=LR, w_latent=1.0, w_pix=1.0,
                                     final_noise_mode='random' if noise is not None else 'const', final_noise=noise,
                                     return_losses=True)
    eng.debug_check()
    return img.cpu(), w_aug.cpu(), losses.cpu()


@pytest.mark.parametrize('name', ['loop_c2.pt', 'loop_c3.pt'])
@pytest.mark.parametrize('precision', ['fp32_parity', 'bf16'])
def test_bench_shape_matches_reference_golden(golden, name, precision):
    from oracle import synthetic
    g = golden(name)
    B, k = g['cfg']['batch'], g['keep_images']
    wl = synthetic.make_workload(g['config'], noise_strength=g['noise_strength'], batch=B)
    eng = _engine(wl, precision, B)
    eng.set_latent_bank(wl['W'])
    eng.set_image_bank(wl['X'])
    ws0 = wl['w0'].repeat(1, wl['G'].num_ws, 1)
    e0 = rel_l2(eng.synthesis(ws0, noise_mode='const').cpu()[:k], g['img0_const'])
    torch.manual_seed(1234)
    noise = [torch.randn([B, 1, r, r]) for r in eng.conv_res]
    img, w_aug, losses = _run(eng, wl, g['steps'], noise)
    ew, ei = rel_l2(w_aug, g['w_aug']), rel_l2(img[:k], g['img'])
    nflip, ew_rest, dmax = flip_stats(w_aug, g['w_aug'])
    # images of the samples WITHOUT a flipped component isolate the forward error from the optimiser's coin tosses
    print(f'\n[golden {name} {precision}] img0={e0:.3e} rel_w={ew:.3e} rel_img={ei:.3e} sign-flips={nflip}/{w_aug.numel()} '
          f'rel_w(no flips)={ew_rest:.3e} max|dw|(no flips)={dmax:.2e}')
    _record(f'{name}:{precision}', img0=e0, rel_w=ew, rel_img=ei, flips=nflip, n=w_aug.numel(), rel_w_noflip=ew_rest, max_dw_noflip=dmax)
    assert e0 < (1e-4 if precision == 'fp32_parity' else 1e-2)
    assert abs(float(losses[0, 0]) - g['loss_latent0']) <= 1e-4 * abs(g['loss_latent0'])
    assert abs(float(losses[0, 1]) - g['loss_pix0']) <= (1e-3 if precision == 'fp32_parity' else 2e-2) * abs(g['loss_pix0'])
    assert ew < TOL[precision] and ei < TOL[precision]
    assert ew_rest < TOL[precision] / (2 if precision == 'fp32_parity' else 1)


@pytest.fixture(scope='module')
def c2_oracle():
    """Oracle loop at C2's own batch (32) and step count (10) on the host cores."""
    from oracle import latent_aug as ola
    from oracle import synthetic
    torch.set_num_threads(os.cpu_count() or 1)
    wl = synthetic.make_workload('c2', noise_strength=0.1)
    orc = ola.LatentAugOracle(wl['G'], wl['W'], wl['X'], num_epochs=10, fused=True)
    random.seed(0)
    _, w_ref = orc.forward(wl['w0'].clone())
    with torch.no_grad():
        img_ref = wl['G'].synthesis(w_ref, noise_mode='const')
    return wl, w_ref[:, 0].contiguous(), img_ref, orc.loss_log


@pytest.mark.parametrize('precision', ['bf16', 'fp32_parity'])
def test_c2_full_batch_matches_oracle(c2_oracle, precision):
    wl, w_ref, img_ref, loss_log = c2_oracle
    eng = _engine(wl, precision, 32)
    eng.set_latent_bank(wl['W'])
    eng.set_image_bank(wl['X'])
    img, w_aug, losses = _run(eng, wl, 10)
    ew, ei = rel_l2(w_aug, w_ref), rel_l2(img, img_ref)
    nflip, ew_rest, dmax = flip_stats(w_aug, w_ref)
    per_sample = [(rel_l2(img[i], img_ref[i])) for i in range(img.shape[0])]
    print(f'\n[C2 B=32 x 10 steps {precision}] rel_w={ew:.3e} rel_img={ei:.3e} (worst sample {max(per_sample):.3e}) '
          f'sign-flips={nflip}/{w_aug.numel()} rel_w(no flips)={ew_rest:.3e} max|dw|(no flips)={dmax:.2e}')
    _record(f'c2_b32:{precision}', rel_w=ew, rel_img=ei, worst_img=max(per_sample), flips=nflip, n=w_aug.numel(),
            rel_w_noflip=ew_rest, max_dw_noflip=dmax)
    for t in (0, 9):
        assert abs(float(losses[t, 0]) - loss_log[t][0]) <= 1e-3 * abs(loss_log[t][0])
        assert abs(float(losses[t, 1]) - loss_log[t][1]) <= (1e-3 if precision == 'fp32_parity' else 2e-2) * abs(loss_log[t][1])
    assert ew < TOL[precision] and ei < TOL[precision]
    assert ew_rest < TOL[precision] / (2 if precision == 'fp32_parity' else 1)


def test_c3_synthesis_and_loop_match_oracle():
    """512x512 generator (64-channel top layers, BN = 64 path): forward and a 2-step loop at batch 2, bf16 and fp32_parity."""
    from oracle import latent_aug as ola
    from oracle import synthetic
    wl = synthetic.make_workload('c3', noise_strength=0.1, batch=2)
    G = wl['G']
    orc = ola.LatentAugOracle(G, wl['W'], wl['X'], num_epochs=2, fused=True)
    random.seed(0)
    _, w_ref = orc.forward(wl['w0'].clone())
    with torch.no_grad():
        img_ref = G.synthesis(w_ref, noise_mode='const')
    for precision in ('fp32_parity', 'bf16'):
        eng = _engine(wl, precision, 2)
        eng.set_latent_bank(wl['W'])
        eng.set_image_bank(wl['X'])
        img, w_aug, _ = _run(eng, wl, 2)
        ew, ei = rel_l2(w_aug, w_ref[:, 0]), rel_l2(img, img_ref)
        nflip, ew_rest, dmax = flip_stats(w_aug, w_ref[:, 0])
        print(f'\n[C3 512^2 B=2 x 2 steps {precision}] rel_w={ew:.3e} rel_img={ei:.3e} sign-flips={nflip} rel_w(no flips)={ew_rest:.3e}')
        _record(f'c3_b2:{precision}', rel_w=ew, rel_img=ei, flips=nflip, n=w_aug.numel(), rel_w_noflip=ew_rest, max_dw_noflip=dmax)
        assert ew < TOL[precision] and ei < TOL[precision]
        del eng
        torch.cuda.empty_cache()
