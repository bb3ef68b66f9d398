"""Oracle restatement of the perceptual (LPIPS) term.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Reference: ``augments/criteria/lpips/networks.py`` (``BaseNet.forward`` :52-64 -- z-score, walk ``vgg16.features``,
tap after 1-based layer i in ``target_layers``, ``normalize_activation``; ``VGG16`` :87-97 taps [16, 23, 30] = relu3_3,
relu4_3, relu5_3, alternative [4, 9, 16, 23, 30] in its comment; ``LinLayers`` :22-32 = 1x1 conv C->1, no bias),
``augments/criteria/lpips/utils.py:6-8`` (``normalize_activation``), ``augments/criteria/lpips/lpips.py:44-68``
(``forward`` / ``forward_tr``: ``sum_layers mean_hw lin((fx - fy)^2)``), and the two loss forms of
``augments/utils/util_latent_aug.py``: ``calc_loss_lpips_torchscript`` :387-409 (mean over all (sample, bank) pairs of the
squared L2 between LPIPS feature vectors) and ``calc_loss_lpips_tr`` :411-424 (sum over pairs / bank size; the
reference's own code broadcasts [B,...] against [M,...] and reads an undefined ``self._modalities`` -- SURVEY.md F10 -- the
pairwise sum is its evident intent and is what is restated).  Crop: ``util_dataset.py:284-332``.

Weights: the pretrained VGG16 / LPIPS-lin / NVIDIA ``vgg16.pt`` files are not available offline; parity is pinned with
seeded random weights run through the REFERENCE's own ``BaseNet`` / ``LinLayers`` / ``LPIPS.forward`` code
(oracle/make_golden_lpips.py -> tests/golden/lpips.pt).  The NVIDIA TorchScript model itself ("lpips_script") cannot be
pinned: it is restated as the same VGG16 with all five taps, whose feature-vector distance is the LPIPS distance.
"""
import math

import torch
import torch.nn.functional as F

VGG_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512]
CONV_IDX = [0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28]      # torchvision vgg16.features module indices of the convs
TAP_LAYERS = (4, 9, 16, 23, 30)          # 1-based indices of BaseNet.forward's enumerate(..., 1): relu1_2 ... relu5_3
TAP_CHANNELS = (64, 128, 256, 512, 512)
TAPS_INTREE = (16, 23, 30)               # networks.py:94
TAPS_SCRIPT = (4, 9, 16, 23, 30)         # NVIDIA vgg16.pt (return_lpips=True) uses all five
MEAN = (-.030, -.088, -.188)             # networks.py:41-44
STD = (.458, .448, .450)


def random_vgg_state(seed=7, taps=TAPS_INTREE):
    """Seeded random weights with torchvision's names (``features.{i}.weight/bias``) + LPIPS lin layers
    (``lin.{k}.weight`` [1, C, 1, 1], k over the used taps in order).  He-scaled so activations stay O(1)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    cin = 3
    ci = 0
    for v in VGG_CFG:
        if v == 'M':
            continue
        i = CONV_IDX[ci]
        sd[f'features.{i}.weight'] = torch.randn([v, cin, 3, 3], generator=g) * math.sqrt(2.0 / (9 * cin))
        sd[f'features.{i}.bias'] = torch.randn([v], generator=g) * 0.05
        cin = v
        ci += 1
    for k, t in enumerate(taps):
        c = TAP_CHANNELS[TAP_LAYERS.index(t)]
        sd[f'lin.{k}.weight'] = (torch.rand([1, c, 1, 1], generator=g) * 0.5 + 0.05)
    return sd


def normalize_activation(x, eps=1e-10):
    """utils.py:6-8"""
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)


def vgg_features(state, x, taps=TAPS_INTREE):
    """networks.py:49-64 -- x [n, 3, h, w] in [-1, 1] -> list of channel-normalised activations at the taps."""
    mean = torch.tensor(MEAN, dtype=x.dtype, device=x.device)[None, :, None, None]
    std = torch.tensor(STD, dtype=x.dtype, device=x.device)[None, :, None, None]
    x = (x - mean) / std
    out = []
    layer = 0          # 1-based index of the module just applied
    ci = 0
    for v in VGG_CFG:
        if v == 'M':
            x = F.max_pool2d(x, 2, 2)
            layer += 1
        else:
            i = CONV_IDX[ci]
            ci += 1
            x = F.conv2d(x, state[f'features.{i}.weight'].to(x), state[f'features.{i}.bias'].to(x), padding=1)
            layer += 1
            x = F.relu(x)
            layer += 1
            if layer in taps:
                out.append(normalize_activation(x))
        if len(out) == len(taps):
            break
    return out


def lin_weights(state, taps=TAPS_INTREE):
    return [state[f'lin.{k}.weight'].reshape(-1) for k in range(len(taps))]


def pair_distance(fx, fy, lw):
    """d[j, i] = sum_layers mean_hw sum_c w_c (fx_i - fy_j)^2  (lpips.py:44-56 for every (sample i, bank j) pair;
    [bank, batch] orientation like l2_loss_vectorized)."""
    d = 0.0
    for a, b, w in zip(fx, fy, lw):
        w = w.to(a).reshape(1, 1, -1, 1, 1)
        diff = (a.unsqueeze(0) - b.unsqueeze(1)) ** 2              # [m, n, C, h, w]
        d = d + (diff * w).sum(2).mean((2, 3))
    return d


def crop(img, pos, size):
    """util_dataset.py:325-332 (applied after the centre crop, :304-309)."""
    x1, y1 = pos
    return img[:, :, y1:y1 + size, x1:x1 + size]


def calc_loss_lpips(state, x_crop, bank_feats, w_lpips, taps=TAPS_INTREE, script=True):
    """x_crop [n, C, s, s]; bank_feats[c] = list over taps of [m, C_l, h_l, w_l] normalised bank activations of
    modality c.  script=True: ULA:387-409 (pair mean); False: ULA:411-424 (pair sum / bank size)."""
    lw = lin_weights(state, taps)
    loss = 0.0
    C = x_crop.shape[1]
    for c in range(C):
        x = x_crop[:, c:c + 1].repeat(1, 3, 1, 1)
        fx = vgg_features(state, x, taps)
        D = pair_distance(fx, bank_feats[c], lw)                   # [m, n]
        m, n = D.shape
        loss_mode = D.sum() / (n * m) if script else D.sum() / m
        loss = loss + loss_mode * w_lpips
    return loss / C


def bank_features(state, bank_crops, taps=TAPS_INTREE):
    """bank_crops [m, C, s, s] (already cropped, in [-1, 1]) -> per modality, per tap normalised activations."""
    out = []
    with torch.no_grad():
        for c in range(bank_crops.shape[1]):
            out.append(vgg_features(state, bank_crops[:, c:c + 1].repeat(1, 3, 1, 1), taps))
    return out
