"""Generate tests/golden/lpips.pt by running the REFERENCE's own LPIPS code (build container only; imports
``/root/reference``).

    python -m oracle.make_golden_lpips [--out tests/golden]

Executed from the reference: ``augments/criteria/lpips/networks.py`` ``BaseNet.__init__/z_score/forward`` (through a
``VGG16`` instance built with ``__new__`` -- its ``__init__`` downloads the torchvision weights, which do not exist
offline -- over a torchvision ``vgg16(weights=None).features`` stack holding seeded random weights), ``LinLayers``,
``utils.normalize_activation``, ``lpips.LPIPS.forward`` (one (sample, bank image) pair per row) and ``util_dataset``'s
``get_params`` / ``get_transform`` crop pipeline.  Autograd through that code gives the gradient pins.
"""
import argparse
import os
import random
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from oracle import lpips as olp
from oracle.ref_driver import REF_SRC, import_reference


def reference_lpips(state, taps):
    """The reference's LPIPS module (criteria/lpips/lpips.py) holding ``state``'s weights."""
    import torchvision
    import_reference(REF_SRC)
    import utils as ref_utils
    if 'utils.util_reports' not in sys.modules:          # networks.py imports it for a commented-out debug plot; it needs bokeh
        stub = types.ModuleType('utils.util_reports')
        sys.modules['utils.util_reports'] = stub
        ref_utils.util_reports = stub
    from augments.criteria.lpips import lpips as rl
    from augments.criteria.lpips import networks as rn
    net = rn.VGG16.__new__(rn.VGG16)
    rn.BaseNet.__init__(net)
    net.report_dir = None
    net.layers = torchvision.models.vgg16(weights=None).features
    net.layers.load_state_dict({k[len('features.'):]: v for k, v in state.items() if k.startswith('features.')})
    net.target_layers = list(taps)
    net.n_channels_list = [olp.TAP_CHANNELS[olp.TAP_LAYERS.index(t)] for t in taps]
    net.set_requires_grad(False)
    m = rl.LPIPS.__new__(rl.LPIPS)
    torch.nn.Module.__init__(m)
    m.net = net
    m.lin = rn.LinLayers(net.n_channels_list)
    m.lin.load_state_dict({f'{k}.1.weight': state[f'lin.{k}.weight'] for k in range(len(taps))})
    m.target_layers = net.target_layers
    return m.eval()


def case(name, taps, script, n=3, m=4, C=2, res=128, size=64, seed=21):
    ref = import_reference(REF_SRC)
    state = olp.random_vgg_state(seed=7, taps=taps)
    M = reference_lpips(state, taps)
    g = torch.Generator().manual_seed(seed)
    img = (torch.rand([n, C, res, res], generator=g) * 2 - 1).requires_grad_(True)
    bank_crops = torch.rand([m, C, size, size], generator=g) * 2 - 1
    random.seed(5)
    params = ref.uds.get_params(load_size=res, crop_size=size, preprocess='center_random_crop')
    tr = ref.uds.get_transform(load_size=res, crop_size=size, preprocess='center_random_crop', params=params)
    x_crop = tr(img)
    w_lpips = 0.7
    loss = 0.0
    feats = []
    for c in range(C):
        x = x_crop[:, c, :, :].unsqueeze(dim=1).repeat([1, 3, 1, 1])            # ULA:395
        y = bank_crops[:, c, :, :].unsqueeze(dim=1).repeat([1, 3, 1, 1])
        D = torch.stack([torch.stack([M(x[i:i + 1], y[j:j + 1]) for i in range(n)]) for j in range(m)])   # [m, n] of LPIPS.forward
        loss_mode = D.sum() / (n * m) if script else D.sum() / m
        loss = loss + loss_mode * w_lpips
        feats.append([f.detach() for f in M.net(x)])
    loss = loss / C
    (grad,) = torch.autograd.grad(loss, img)
    return dict(name=name, taps=tuple(taps), script=script, state=state, img=img.detach(), bank_crops=bank_crops, crop_pos=params['crop_pos'],
                crop_size=size, res=res, w_lpips=w_lpips, loss=float(loss), grad=grad, feats=feats)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='tests/golden')
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    out = {'intree3': case('intree3', olp.TAPS_INTREE, False), 'script5': case('script5', olp.TAPS_SCRIPT, True, n=2, m=3)}
    # keep the file small: features of sample 0 / modality 0 of the 3-tap case, per-tap sums of the other
    for k, c in out.items():
        c['feat_sums'] = [[(float(f.sum()), float(f.abs().sum())) for f in mode] for mode in c['feats']]
        c['feats'] = [f[:1].clone() for f in c['feats'][0]] if k == 'intree3' else None
        c['state'] = None                 # rebuilt from the seed (oracle.lpips.random_vgg_state(7, taps))
    torch.save(out, os.path.join(args.out, 'lpips.pt'))
    print('lpips.pt', os.path.getsize(os.path.join(args.out, 'lpips.pt')))


if __name__ == '__main__':
    main()
