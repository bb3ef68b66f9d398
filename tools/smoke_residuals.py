"""Residual of the smoke() workload against the CPU oracle as a function of the number of optimisation steps
(0 = synthesis only).  Separates forward-path error from the optimiser's amplification of rounding-order changes."""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from latentaugment_b200.engine import SynthesisEngine
from oracle import latent_aug as ola
from oracle import synthetic


def rel(a, b):
    return float((a.double().cpu() - b.double()).norm() / b.double().norm())


wl = synthetic.make_workload('tiny', noise_strength=0.1)
G = wl['G']
for precision in ('fp32_parity', 'bf16'):
    eng = SynthesisEngine(dict(G.state_dict()), img_resolution=G.img_resolution, img_channels=G.img_channels, w_dim=G.w_dim,
                          z_dim=G.z_dim, batch=wl['w0'].shape[0], precision=precision, device='cuda:0')
    eng.set_latent_bank(wl['W'])
    eng.set_image_bank(wl['X'])
    for steps in (0, 1, 2, 3, 6):
        orc = ola.LatentAugOracle(G, wl['W'], wl['X'], num_epochs=steps)
        random.seed(0)
        _, w_ref = orc.forward(wl['w0'].clone())
        img, w_aug = eng.augment(wl['w0'], num_steps=steps, lr=0.01, final_noise_mode='const')
        torch.cuda.synchronize()
        with torch.no_grad():
            img_ref = G.synthesis(w_ref, noise_mode='const')
        # the same image synthesised by the ENGINE from the ORACLE's final w: forward-path error alone
        img_fwd = eng.synthesis(w_ref, noise_mode='const') if hasattr(eng, 'synthesis') else None
        line = f'{precision:12s} steps={steps}: rel_l2 w={rel(w_aug, w_ref[:, 0]):.3e} img={rel(img, img_ref):.3e}'
        if img_fwd is not None:
            line += f'  forward-only img={rel(img_fwd, img_ref):.3e}'
        print(line, flush=True)
