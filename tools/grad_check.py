"""Per-layer gradient check at a benchmark shape: style gradients and d loss / d w of ONE optimisation step of the CUDA
path (la_debug_get) against autograd through the CPU oracle, layer by layer -- localises a wrong backward launch.

    python tools/grad_check.py --config c2 --batch 8 [--precision fp32_parity] [--simt] [--cache /tmp/gc.pt]

Environment switches of the kernels (LA_CTA2, LA_BN, LA_NO_STAGED, LA_SEED_STREAM, LA_NO_FIR_TMA, LA_UPCONV_SPLIT_MIN_RES,
LA_NO_GRAPH) are read by the library as usual, so the same command under different switches shows which path is off.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch


def oracle_grads(wl, cache):
    if cache and os.path.exists(cache):
        return torch.load(cache, weights_only=True)
    from oracle import latent_aug as ola
    G = wl['G']
    torch.set_num_threads(os.cpu_count() or 1)
    names, mods = [], []
    S = G.synthesis
    for r in S.block_resolutions:
        blk = getattr(S, f'b{r}')
        if blk.cin != 0:
            names.append(f'b{r}.conv0'); mods.append(blk.conv0)
        names.append(f'b{r}.conv1'); mods.append(blk.conv1)
    for r in S.block_resolutions:
        names.append(f'b{r}.torgb'); mods.append(getattr(S, f'b{r}').torgb)
    outs = {}
    hooks = []
    for nm, m in zip(names, mods):
        def hook(mod, inp, out, nm=nm):
            out.retain_grad()
            outs[nm] = out
        hooks.append(m.affine.register_forward_hook(hook))
    w = wl['w0'].clone().requires_grad_(True)
    ws = w.repeat(1, G.num_ws, 1)
    x = G.synthesis(ws, noise_mode='const', fused=True)
    res = G.img_resolution
    l_pix = ola.calc_loss_pix(ola.center_crop(x, res), ola.center_crop(wl['X'], res), 1.0, G.img_channels)
    (-l_pix).backward()
    for h in hooks:
        h.remove()
    out = dict(names=names, g_s={k: v.grad.detach().clone() for k, v in outs.items()}, s={k: v.detach().clone() for k, v in outs.items()},
               grad_w=w.grad.detach()[:, 0].clone(), l_pix=float(l_pix), img=x.detach())
    if cache:
        torch.save(out, cache)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--config', default='c2')
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--precision', default='fp32_parity')
    ap.add_argument('--simt', action='store_true')
    ap.add_argument('--cache', default='')
    ap.add_argument('--tag', default='')
    ap.add_argument('--steps', type=int, default=1, help='> 1: only compare the final w of a k-step loop with the oracle loop (sign flips)')
    a = ap.parse_args()
    from latentaugment_b200.engine import SynthesisEngine
    from oracle import synthetic
    wl = synthetic.make_workload(a.config, noise_strength=0.1, batch=a.batch)
    G = wl['G']
    ref = oracle_grads(wl, a.cache) if a.steps == 1 else None
    eng = SynthesisEngine(dict(G.state_dict()), img_resolution=G.img_resolution, img_channels=G.img_channels, w_dim=G.w_dim, z_dim=G.z_dim,
                          batch=a.batch, precision=a.precision)
    eng.set_latent_bank(wl['W'])
    eng.set_image_bank(wl['X'])
    if a.simt:
        eng.debug_set_simt(1)
    if a.steps > 1:
        import random

        from oracle import latent_aug as ola
        cache = (a.cache or '/tmp/gc') + f'.loop{a.steps}.pt'
        if os.path.exists(cache):
            w_ref = torch.load(cache)
        else:
            torch.set_num_threads(os.cpu_count() or 1)
            orc = ola.LatentAugOracle(G, wl['W'], wl['X'], num_epochs=a.steps, fused=True)
            random.seed(0)
            w_ref = orc.forward(wl['w0'].clone())[1][:, 0].contiguous()
            torch.save(w_ref, cache)
        _, w_aug = eng.augment(wl['w0'], num_steps=a.steps, lr=0.01, final_noise_mode='const')
        eng.debug_check()
        d = (w_aug.cpu().double() - w_ref.double()).abs()
        tag = a.tag or ' '.join(f'{k}={v}' for k, v in os.environ.items() if k.startswith('LA_')) or 'default'
        print(f'== loop {a.config} B={a.batch} {a.precision} steps={a.steps} simt={int(a.simt)} [{tag}]: rel_w={float(d.norm() / w_ref.double().norm()):.3e} '
              f'components off by > lr: {int((d > 0.01).sum())}/{d.numel()}  > lr/10: {int((d > 0.001).sum())}  max {float(d.max()):.4f}')
        return
    img, w_aug, losses = eng.augment(wl['w0'], num_steps=1, lr=0.01, w_latent=0.0, w_pix=1.0, final_noise_mode='const', return_losses=True)
    eng.debug_check()
    g_s, s, gw = eng.debug_get('g_s').cpu(), eng.debug_get('s').cpu(), eng.debug_get('grad_w').cpu().reshape(a.batch, -1)
    soff = [int(v) for v in eng.debug_get('soff').cpu().tolist()]
    B = a.batch

    def rel(x, y):
        return float((x.double() - y.double()).norm() / y.double().norm().clamp_min(1e-30))

    def cos(x, y):
        return float(torch.nn.functional.cosine_similarity(x.double().flatten(), y.double().flatten(), dim=0))
    tag = a.tag or ' '.join(f'{k}={v}' for k, v in os.environ.items() if k.startswith('LA_')) or 'default'
    print(f'== grad_check {a.config} B={B} {a.precision} simt={int(a.simt)} [{tag}]  l_pix ours={float(losses[0, 1]):.6f} oracle={ref["l_pix"]:.6f}')
    sign = int((torch.sign(gw) != torch.sign(ref['grad_w'])).sum())
    print(f'   d loss/d w: rel={rel(gw, ref["grad_w"]):.3e} cos={cos(gw, ref["grad_w"]):.6f} sign disagreements {sign}/{gw.numel()}')
    for nm, off in zip(ref['names'], soff):
        go, so = ref['g_s'][nm], ref['s'][nm]
        cin = go.shape[1]
        mine = g_s[B * off:B * off + B * cin].reshape(B, cin)
        smine = s[B * off:B * off + B * cin].reshape(B, cin)
        ratio = float(mine.norm() / go.norm().clamp_min(1e-30))
        print(f'   {nm:12s} cin={cin:4d}  style rel={rel(smine * float(so.norm() / smine.norm()), so):.2e}  g_s cos={cos(mine, go):.6f} |ours|/|oracle|={ratio:.4f} '
              f'rel(after rescale)={rel(mine / max(ratio, 1e-30), go):.3e}')


if __name__ == '__main__':
    main()
