"""Small end-to-end case for compute-sanitizer: tiny generator + discriminator loop (bf16 and fp32_parity),
a 64^2 generator (TMA FIR passes, bulk seed kernel), nearest codes and filtered_lrelu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from latentaugment_b200.engine import LatentBank, SynthesisEngine
from latentaugment_b200.ops_sg3 import filtered_lrelu
from oracle import sg2_disc, synthetic

for cfg, precs in (('tiny', ('bf16', 'fp32_parity')),
                   (dict(img_resolution=64, img_channels=3, channel_base=8192, channel_max=128, batch=3, steps=2, bank=32, img_bank=4), ('bf16',))):
    wl = synthetic.make_workload(cfg, noise_strength=0.1)
    G, c = wl['G'], wl['cfg']
    D = sg2_disc.make_discriminator(img_resolution=c['img_resolution'], img_channels=c['img_channels'], channel_base=c['channel_base'],
                                    channel_max=c['channel_max'])
    for prec in precs:
        B = wl['w0'].shape[0]
        eng = SynthesisEngine(dict(G.state_dict()), img_resolution=G.img_resolution, img_channels=G.img_channels, w_dim=G.w_dim, z_dim=G.z_dim,
                              batch=B, precision=prec)
        eng.set_latent_bank(wl['W'])
        eng.set_image_bank(wl['X'])
        if B % 4 == 0:
            eng.set_discriminator(dict(D.state_dict()))
        for it in range(2):
            img, w = eng.augment(wl['w0'], num_steps=3, w_disc=1.0 if B % 4 == 0 else 0.0, final_noise_mode='random')
        torch.cuda.synchronize()
        eng.debug_check()
        print(cfg if isinstance(cfg, str) else 'res64', prec, float(img.abs().mean()))
X = torch.randn(50, 512).cuda()
Y = torch.randn(1000, 512).cuda()
print(LatentBank(Y).nearest(X, 4)[1][:2])
x = torch.randn(2, 5, 20, 20, device='cuda', requires_grad=True)
f = torch.rand(12) + 0.1
y = filtered_lrelu(x, f / f.sum(), f / f.sum(), torch.randn(5, device='cuda'), up=2, down=2, padding=10, clamp=1.0)
y.sum().backward()
torch.cuda.synchronize()
print('ok', y.shape, float(x.grad.abs().mean()))
