"""Generate tests/golden/*.pt by running the REFERENCE's own code (build container only).

Usage:  python -m oracle.make_golden [--out tests/golden] [--skip-c1]

Imports ``/root/reference`` (read-only; it does not exist on the GPU box, which is why
the outputs are committed).  What is executed from the reference:
  * ``torch_utils.ops``: ``_bias_act_ref``, ``_upfirdn2d_ref``, ``upsample2d``,
    ``setup_filter``, ``conv2d_resample``, ``fma``  (op-level pins)
  * ``LatentAug.l2_loss_vectorized / calc_loss_latent / calc_loss_pix``  (loss-level pins)
  * ``LatentAug.forward`` itself -- the 10-step Adam loop -- constructed via
    ``LatentAug.__new__`` (its ``__init__`` needs zips/pickles that do not exist offline,
    SURVEY.md §8c) and driven through ``oracle.ref_driver.RefOpsGenerator``: the oracle generator's
    parameters, but every op call routed to the reference's ``torch_utils.ops``.
The generator *classes* are not in the reference, so the layer
composition in ``RefOpsGenerator`` is this repo's restatement; ops, losses and the loop
are the reference's.
"""
import argparse
import hashlib
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from oracle.ref_driver import REF_SRC, RefOpsGenerator, import_reference, make_reference_latentaug   # noqa: E402


def _import_reference():
    return import_reference(REF_SRC)


def digest(*tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]


def op_goldens(ref):
    g = torch.Generator().manual_seed(11)
    out = {}
    x = torch.randn([2, 6, 9, 7], generator=g)
    b = torch.randn([6], generator=g)
    cases = []
    for act, gain, clamp in [('lrelu', None, 256.0), ('lrelu', 0.7, 0.5), ('linear', None, 1.0), ('linear', None, None),
                             ('relu', None, None), ('sigmoid', None, None), ('swish', 2.0, 1.5)]:
        xr = (x * 3).clone().requires_grad_(True)
        y = ref.bias_act._bias_act_ref(xr, b, act=act, gain=gain, clamp=clamp)
        dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5))
        (dx,) = torch.autograd.grad(y, xr, dy)
        cases.append(dict(act=act, gain=gain, clamp=clamp, y=y.detach(), dx=dx))
    out['bias_act'] = dict(x=x * 3, b=b, cases=cases)

    f = ref.upfirdn2d.setup_filter([1, 3, 3, 1])
    out['setup_filter_1331'] = f
    out['setup_filter_sep12'] = ref.upfirdn2d.setup_filter(list(range(1, 13)))
    xs = torch.randn([2, 3, 8, 6], generator=g)
    ucases = []
    for kw in [dict(up=2, padding=[2, 1, 2, 1], gain=4.0), dict(up=1, padding=[1, 1, 1, 1], gain=4.0),
               dict(down=2, padding=[1, 1, 1, 1]), dict(up=2, down=1, padding=[-1, 2, 0, 3], flip_filter=True),
               dict(up=[2, 1], down=[1, 2], padding=[1, 2, 2, 1], gain=2.0)]:
        xr = xs.clone().requires_grad_(True)
        y = ref.upfirdn2d._upfirdn2d_ref(xr, f, **kw)
        dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(6))
        (dx,) = torch.autograd.grad(y, xr, dy)
        ucases.append(dict(kw=kw, y=y.detach(), dx=dx))
    out['upfirdn2d'] = dict(x=xs, f=f, cases=ucases)
    out['upsample2d'] = dict(x=xs, y=ref.upfirdn2d.upsample2d(xs, f, impl='ref'))

    xc = torch.randn([2, 4, 6, 5], generator=g)
    ccases = []
    for cout, k, kw in [(5, 3, dict(up=2, padding=1, flip_weight=False)), (5, 3, dict(padding=1)),
                        (3, 1, dict()), (5, 3, dict(down=2, padding=1)), (3, 1, dict(up=2)),
                        (3, 1, dict(down=2)), (4, 3, dict(up=2, padding=1, groups=2, flip_weight=False))]:
        groups = kw.get('groups', 1)
        w = torch.randn([cout, 4 // groups, k, k], generator=g)
        xr = xc.clone().requires_grad_(True)
        y = ref.conv2d_resample.conv2d_resample(xr, w, f=(f if (kw.get('up', 1) > 1 or kw.get('down', 1) > 1) else None), **kw)
        dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(7))
        (dx,) = torch.autograd.grad(y, xr, dy)
        ccases.append(dict(kw=kw, w=w, y=y.detach(), dx=dx))
    out['conv2d_resample'] = dict(x=xc, f=f, cases=ccases)

    a, bb, c = torch.randn([2, 3, 4, 4], generator=g), torch.randn([2, 3, 1, 1], generator=g), torch.randn([1, 1, 4, 4], generator=g)
    out['fma'] = dict(a=a, b=bb, c=c, y=ref.fma.fma(a, bb, c))
    return out


def loss_goldens(ref):
    g = torch.Generator().manual_seed(12)
    L = ref.ula.LatentAug
    out = {}
    for nd, (xs, ys) in {2: ([5, 33], [7, 33]), 3: ([4, 6, 16], [9, 6, 16]), 4: ([3, 1, 9, 9], [6, 1, 9, 9])}.items():
        X, Y = torch.randn(xs, generator=g), torch.randn(ys, generator=g) * 1.5 + 0.3
        out[f'l2_{nd}d'] = dict(X=X, Y=Y, D=L.l2_loss_vectorized(X, Y, compute_mean=False),
                                mean=L.l2_loss_vectorized(X, Y, compute_mean=True))
    return out


def loop_golden(ref, name, noise_strength, steps=None, w_latent=1.0, w_pix=1.0, soft_aug=False, alpha=1.0, batch=None,
                keep_images=None):
    """``batch`` overrides the config's batch size; ``keep_images`` stores the images of the first k samples only
    (the benchmark shapes: a full C2 / C3 image batch is 25 / 400 MB)."""
    from oracle import synthetic
    wl = synthetic.make_workload(name, noise_strength=noise_strength, batch=batch)
    cfg = wl['cfg']
    if batch is not None:
        cfg['batch'] = batch
    steps = cfg['steps'] if steps is None else steps
    Gref = RefOpsGenerator(wl['G'], ref)
    la = make_reference_latentaug(ref, Gref, wl['W'], wl['X'], cfg, steps, w_latent, w_pix, soft_aug, alpha)
    # loss-level pins on real shapes
    with torch.no_grad():
        ws0 = la.broadcasting(wl['w0'])
        x0 = Gref.synthesis(ws0, noise_mode='const')
        ccrop = __import__('augments.utils.util_dataset', fromlist=['x']).get_center_crop(load_size=cfg['img_resolution'])
        l_lat0 = float(la.calc_loss_latent(ws0, la.W))
        l_pix0 = float(la.calc_loss_pix(ccrop(x0), ccrop(la.X)))
        D0 = la.l2_loss_vectorized(ws0, la.W, compute_mean=False)
    random.seed(0)
    torch.manual_seed(1234)
    img, w_aug = la(wl['w0'].clone(), ['synthetic'] * cfg['batch'])
    if keep_images is not None:
        x0, img = x0[:keep_images].clone(), img[:keep_images].clone()
    return dict(config=name, cfg=cfg, keep_images=keep_images, noise_strength=noise_strength, steps=steps, w_latent=w_latent, w_pix=w_pix,
                soft_aug=soft_aug, alpha=alpha, inputs_digest=digest(wl['W'], wl['w0'], wl['X']),
                params_digest=digest(*[p for p in wl['G'].state_dict().values()]),
                img0_const=x0, loss_latent0=l_lat0, loss_pix0=l_pix0,
                nn_idx0=D0.argmin(0), nn_top4=D0.topk(min(4, D0.shape[0]), dim=0, largest=False).indices.t().contiguous(),
                img=img.detach(), w_aug=w_aug.detach()[:, 0, :].contiguous())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='tests/golden')
    ap.add_argument('--skip-c1', action='store_true')
    ap.add_argument('--only', default='', help='comma list of fixture names to (re)generate, e.g. loop_c2,loop_c3')
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    ref = _import_reference()
    os.makedirs(args.out, exist_ok=True)
    big = {   # the benchmarked shapes: C2 at batch 8 x 10 steps, C3 (512^2) at batch 2 x 2 steps
        'loop_c2': lambda: loop_golden(ref, 'c2', 0.1, batch=8, keep_images=2),
        'loop_c3': lambda: loop_golden(ref, 'c3', 0.1, steps=2, batch=2, keep_images=1),
    }
    if args.only:
        for nm in args.only.split(','):
            torch.save(big[nm](), os.path.join(args.out, nm + '.pt'))
            print(nm, os.path.getsize(os.path.join(args.out, nm + '.pt')))
        return
    torch.save(op_goldens(ref), os.path.join(args.out, 'ops.pt'))
    torch.save(loss_goldens(ref), os.path.join(args.out, 'losses.pt'))
    torch.save(loop_golden(ref, 'tiny', 0.1), os.path.join(args.out, 'loop_tiny.pt'))
    torch.save(loop_golden(ref, 'tiny', 0.1, soft_aug=True, alpha=0.7, w_pix=0.1), os.path.join(args.out, 'loop_tiny_soft.pt'))
    torch.save(loop_golden(ref, 'tiny128', 0.1), os.path.join(args.out, 'loop_tiny128.pt'))
    torch.save(loop_golden(ref, 'small', 0.1), os.path.join(args.out, 'loop_small.pt'))
    if not args.skip_c1:
        torch.save(loop_golden(ref, 'c1', 0.1), os.path.join(args.out, 'loop_c1.pt'))
    for fn in sorted(os.listdir(args.out)):
        print(fn, os.path.getsize(os.path.join(args.out, fn)))


if __name__ == '__main__':
    main()
