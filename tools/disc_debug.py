import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latentaugment_b200.engine import SynthesisEngine
from oracle import sg2_disc, synthetic

cfg = sys.argv[1] if len(sys.argv) > 1 else 'tiny'
wl = synthetic.make_workload(cfg, noise_strength=0.1)
G, c = wl['G'], wl['cfg']
D = sg2_disc.make_discriminator(img_resolution=c['img_resolution'], img_channels=c['img_channels'], channel_base=c['channel_base'], channel_max=c['channel_max'])
eng = SynthesisEngine(dict(G.state_dict()), img_resolution=G.img_resolution, img_channels=G.img_channels, w_dim=G.w_dim, z_dim=G.z_dim, batch=wl['w0'].shape[0], precision='bf16')
eng.set_discriminator(dict(D.state_dict()))
B, R = wl['w0'].shape[0], c['img_resolution']
x = (torch.rand([B, c['img_channels'], R, R], generator=torch.Generator().manual_seed(11)) * 2 - 1).requires_grad_(True)
lg = D(x, c=None)
(torch.nn.functional.softplus(-lg).mean()).backward()
loss, grad = eng.disc_loss_grad(x, w_disc=1.0)
g, r = grad.cpu().double(), x.grad.double()
def rel(a, b): return float((a - b).norm() / b.norm())
print('all', rel(g, r), 'cos', float((g * r).sum() / g.norm() / r.norm()), 'norm ratio', float(g.norm() / r.norm()))
print('interior', rel(g[:, :, 4:-4, 4:-4], r[:, :, 4:-4, 4:-4]), 'border rows', rel(g[:, :, :2], r[:, :, :2]), rel(g[:, :, -2:], r[:, :, -2:]),
      'border cols', rel(g[:, :, :, :2], r[:, :, :, :2]), rel(g[:, :, :, -2:], r[:, :, :, -2:]))
for n in range(B): print('sample', n, rel(g[n], r[n]))
for py in range(2):
    for px in range(2): print('phase', py, px, rel(g[:, :, py::2, px::2], r[:, :, py::2, px::2]))
