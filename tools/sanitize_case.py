"""Small end-to-end case for compute-sanitizer: tiny generator + discriminator loop (bf16 and fp32_parity),
a 64^2 generator (TMA FIR passes, bulk seed kernel), the four-term loop with the perceptual term, nearest codes (incl. the
rescan paths of the re-rank) and filtered_lrelu."""
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
# perceptual term: 128x128 so the 64x64 window fits the centre crop; all four criteria in the loop, both tap sets
import random

from latentaugment_b200.augments.utils.util_latent_aug import feature_bank_crops
from oracle import lpips as olp
cfg = dict(img_resolution=128, img_channels=2, channel_base=8192, channel_max=64, batch=4, steps=2, bank=32, img_bank=6)
wl = synthetic.make_workload(cfg, noise_strength=0.1)
G = wl['G']
D = sg2_disc.make_discriminator(img_resolution=128, img_channels=2, channel_base=8192, channel_max=64)
random.seed(1)
crops = feature_bank_crops(wl['X'], 128, 64)
for prec, taps in (('bf16', olp.TAPS_SCRIPT), ('fp32_parity', olp.TAPS_INTREE)):
    eng = SynthesisEngine(dict(G.state_dict()), img_resolution=128, img_channels=2, batch=4, precision=prec)
    eng.set_latent_bank(wl['W'])
    eng.set_image_bank(wl['X'])
    eng.set_discriminator(dict(D.state_dict()))
    eng.set_lpips(olp.random_vgg_state(7, taps), taps=taps, crop_size=64)
    eng.set_feature_bank(crops)
    for it in range(2):
        img, w = eng.augment(wl['w0'], num_steps=2, w_disc=0.01, w_lpips=10.0, w_pix=0.1, w_latent=0.001, lpips_crop=(3, 17), final_noise_mode='const')
    torch.cuda.synchronize()
    eng.debug_check()
    print('four terms', prec, float(img.abs().mean()))
X = torch.randn(50, 512).cuda()
Y = torch.randn(1000, 512).cuda()
print(LatentBank(Y).nearest(X, 4)[1][:2])
Y[300:340] = X[3] + 1e-3 * torch.randn(40, 512).cuda()          # overflowing chunks -> exhaustive rescans
print(LatentBank(Y).nearest(X, 8)[1][3])
Y = Y[:1].repeat(1000, 1) + 1e-3 * torch.randn(1000, 512).cuda()  # every chunk overflows -> whole-shard scan
print(LatentBank(Y).nearest(X, 8)[1][0])
x = torch.randn(2, 5, 20, 20, device='cuda', requires_grad=True)
f = torch.rand(12) + 0.1
y = filtered_lrelu(x, f / f.sum(), f / f.sum(), torch.randn(5, device='cuda'), up=2, down=2, padding=10, clamp=1.0)
y.sum().backward()
torch.cuda.synchronize()
print('ok', y.shape, float(x.grad.abs().mean()))
