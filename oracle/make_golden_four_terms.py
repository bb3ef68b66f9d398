"""Generate tests/golden/loop_four_terms.pt by running the REFERENCE's own ``LatentAug.forward`` with ALL FOUR criteria
(build container only; imports ``/root/reference``).

    python -m oracle.make_golden_four_terms [--out tests/golden]

Executed from the reference: ``LatentAug.forward`` (util_latent_aug.py:207-310: crop parameters drawn once per call,
``get_transform`` / ``get_center_crop`` pipelines, the four criteria and their sign combination :270, Adam),
``calc_loss_latent`` :427-433, ``calc_loss_pix`` :373-385, ``calc_loss_disc`` :363-371 and
``calc_loss_lpips_torchscript`` :387-409 (``l2_loss_vectorized`` between LPIPS feature VECTORS, mean over all (sample,
bank) pairs), over the reference's ``torch_utils.ops`` (``RefOpsGenerator``).  Weights: the author's
(backbone_latentaug.py:46-49).  What is NOT the reference's because it is not in its tree: the generator and
discriminator classes (restated, oracle/sg2.py, sg2_disc.py) and the NVIDIA TorchScript ``vgg16.pt`` behind
``self.vgg16(x, resize_images=False, return_lpips=True)`` -- stood in for by ``ScriptVGG`` below, the restated VGG16 whose
returned feature vector is built so that the squared L2 distance of two vectors is the LPIPS distance.
"""
import argparse
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from oracle import lpips as olp
from oracle.ref_driver import RefOpsGenerator, import_reference, make_reference_latentaug

CFG = dict(img_resolution=128, img_channels=2, channel_base=8192, channel_max=64, batch=4, steps=3, bank=32, img_bank=6)
WEIGHTS = dict(w_latent=0.001, w_pix=0.1, w_lpips=10.0, w_disc=0.01)


class ScriptVGG(torch.nn.Module):
    """``vgg16(x, resize_images=False, return_lpips=True) -> [n, F]`` with ``|f(x) - f(y)|^2`` = the LPIPS distance:
    per tap ``n^ * sqrt(w_c / (h w))`` flattened, taps concatenated."""

    def __init__(self, state, taps):
        super().__init__()
        self.state, self.taps = state, taps

    def forward(self, x, resize_images=False, return_lpips=True):
        assert not resize_images and return_lpips
        feats = olp.vgg_features(self.state, x, self.taps)
        lw = olp.lin_weights(self.state, self.taps)
        out = []
        for f, w in zip(feats, lw):
            hw = f.shape[2] * f.shape[3]
            out.append((f * torch.sqrt(w.to(f) / hw).reshape(1, -1, 1, 1)).flatten(1))
        return torch.cat(out, dim=1)


def bank_crops(X, res, size):
    """One fresh window per (modality, image) in the order the reference builds ``fea_{mode}`` (util_latent_aug.py:160-169,
    564-579) -- through the reference's own get_params / get_transform."""
    from augments.utils import util_dataset as ud
    out = torch.empty([X.shape[0], X.shape[1], size, size])
    for c in range(X.shape[1]):
        for j in range(X.shape[0]):
            params = ud.get_params(load_size=res, crop_size=size, preprocess='center_random_crop')
            tr = ud.get_transform(load_size=res, crop_size=size, preprocess='center_random_crop', params=params)
            out[j, c] = tr(X[j:j + 1, c:c + 1])[0, 0]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='tests/golden')
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    ref = import_reference()
    from oracle import sg2_disc, synthetic
    wl = synthetic.make_workload(CFG, noise_strength=0.1)
    D = sg2_disc.make_discriminator(img_resolution=128, img_channels=2, channel_base=8192, channel_max=64)
    taps = olp.TAPS_SCRIPT
    st = olp.random_vgg_state(7, taps)
    random.seed(11)
    crops = bank_crops(wl['X'], 128, 64)
    vgg = ScriptVGG(st, taps)
    la = make_reference_latentaug(ref, RefOpsGenerator(wl['G'], ref), wl['W'], wl['X'], dict(CFG), CFG['steps'], WEIGHTS['w_latent'],
                                  WEIGHTS['w_pix'], w_lpips=WEIGHTS['w_lpips'], w_disc=WEIGHTS['w_disc'])
    la.D, la.vgg16 = D, vgg
    with torch.no_grad():
        for c, mode in enumerate(la.modalities):
            la.register_buffer(f'fea_{mode}', vgg(crops[:, c:c + 1].repeat(1, 3, 1, 1)))
    # per-step loss values: the reference only keeps them in a local dict, so record them through its own criteria
    log = []
    hooks = {}
    for name in ('calc_loss_latent', 'calc_loss_pix', 'calc_loss_disc', 'calc_loss_lpips_torchscript'):
        fn = getattr(la, name)

        def wrapped(*args, _fn=fn, _name=name, **kw):
            v = _fn(*args, **kw)
            log.append((_name, float(v.detach())))
            return v
        hooks[name] = wrapped
        setattr(la, name, wrapped)
    random.seed(0)
    torch.manual_seed(1234)
    img, w_aug = la(wl['w0'].clone(), ['synthetic'] * CFG['batch'])
    steps = []
    for t in range(CFG['steps']):
        d = dict(log[4 * t:4 * t + 4])
        steps.append((d['calc_loss_latent'], d['calc_loss_pix'], d['calc_loss_disc'], d['calc_loss_lpips_torchscript']))
    out = dict(cfg=dict(CFG), weights=dict(WEIGHTS), taps=taps, vgg_seed=7, crop_seed=11, loop_seed=0, bank_crops=crops,
               losses=torch.tensor(steps, dtype=torch.float64), img=img.detach(), w_aug=w_aug.detach()[:, 0, :].contiguous())
    os.makedirs(a.out, exist_ok=True)
    path = os.path.join(a.out, 'loop_four_terms.pt')
    torch.save(out, path)
    print(path, os.path.getsize(path), 'losses per step (latent, pix, disc, lpips):', steps)


if __name__ == '__main__':
    main()
