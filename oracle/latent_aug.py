"""Oracle restatement of the latent-optimisation loop and its criteria.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Reference file:
``/root/reference/augments/utils/util_latent_aug.py`` (cited as ULA below) and
``augments/utils/util_dataset.py`` (UDS).
"""
import math
import random

import torch


def l2_loss_vectorized(X, Y, compute_mean=True):
    """Pairwise squared L2 between batch ``X`` [n,...] and bank ``Y`` [m,...].
    Follows ULA:315-361 -- output orientation ``[bank, batch]``; association
    ``(YY[:,None] + XX) - 2*YX``; mean as two separate divisions."""
    if X.ndim not in (2, 3, 4) or Y.ndim != X.ndim:
        raise NotImplementedError
    m, n = Y.shape[0], X.shape[0]
    red = tuple(range(1, X.ndim))
    yy = Y.square().sum(red)
    xx = X.square().sum(red)
    yx = torch.einsum('nk,mk->nm', Y.reshape(m, -1), X.reshape(n, -1))
    D = (yy.unsqueeze(-1) + xx) - 2 * yx
    if compute_mean:
        D = D.sum() / (m * n)
        D = D / math.prod(Y.shape[1:])
    return D


def center_crop_bounds(res):
    """torchvision ``CenterCrop(int(sqrt(res^2/2)))`` as used by UDS:317-323:
    returns (offset, size); offset = int(round((res-size)/2.0))."""
    size = int(math.sqrt((res * res) / 2))
    off = int(round((res - size) / 2.0))
    return off, size


def center_crop(img, res):
    off, size = center_crop_bounds(res)
    return img[:, :, off:off + size, off:off + size]


def get_crop_params(load_size, crop_size, preprocess='center_random_crop'):
    """UDS:284-296 -- consumes two python ``random.randint`` draws per call."""
    new = load_size
    if preprocess == 'center_random_crop':
        new = int(math.sqrt((load_size * load_size) / 2))
    x = random.randint(0, max(0, new - crop_size))
    y = random.randint(0, max(0, new - crop_size))
    return x, y


def calc_loss_latent(ws, W, w_latent):
    """ULA:427-433."""
    return l2_loss_vectorized(ws, W) * w_latent


def calc_loss_pix(x, x_bank, w_pix, n_modalities):
    """ULA:373-385: per-modality mean pairwise L2, weights applied per modality,
    averaged over modalities."""
    loss = 0.0
    for c in range(n_modalities):
        loss = loss + l2_loss_vectorized(x[:, c:c + 1], x_bank[:, c:c + 1]) * w_pix
    return loss / n_modalities


def nearest_codes(ws, W, k=1):
    """The north_star's nearest-code extension: indices of the k bank
    rows closest to each sample, defined on the reference's own distance matrix
    ``D = l2_loss_vectorized(ws, W, compute_mean=False)`` ([bank, batch], ULA:332-340).
    Ties break to the lowest index.  Returns (dist [batch,k], idx [batch,k])."""
    D = l2_loss_vectorized(ws, W, compute_mean=False)        # [M, B]
    d, i = torch.sort(D.t(), dim=1, stable=True)
    return d[:, :k].contiguous(), i[:, :k].contiguous()


class LatentAugOracle:
    """The N-step Adam loop on w (ULA:207-310) with the latent and pixel criteria.

    ``G`` exposes ``.synthesis(ws, noise_mode=...)``, ``.mapping``, ``num_ws``,
    ``w_dim``; ``D`` (optional) is the realism-term discriminator (oracle/sg2_disc.py); ``lpips`` (optional) =
    dict(state, taps, script, bank_feats) configures the perceptual term (oracle/lpips.py).
    """

    def __init__(self, G, W=None, X=None, *, num_epochs=10, opt_lr=0.01, w_latent=1.0, w_pix=1.0,
                 w_lpips=0.0, w_disc=0.0, soft_aug=False, alpha=1.0, truncation_psi=1.0,
                 n_modalities=None, res=None, crop_size=64, preprocess='center_random_crop', fused=True, D=None, lpips=None):
        assert w_lpips == 0.0 or lpips is not None, 'w_lpips > 0 needs the VGG / bank-feature configuration'
        assert w_disc == 0.0 or D is not None, 'w_disc > 0 needs a discriminator'
        self.G, self.W, self.X, self.D, self.w_disc = G, W, X, D, w_disc
        self.w_lpips, self.lpips = w_lpips, lpips
        self.num_ws, self.w_dim = G.num_ws, G.w_dim
        self.num_epochs, self.opt_lr = num_epochs, opt_lr
        self.w_latent, self.w_pix = w_latent, w_pix
        self.soft_aug, self.alpha, self.truncation_psi = soft_aug, alpha, truncation_psi
        self.res = res if res is not None else G.img_resolution
        self.n_modalities = n_modalities if n_modalities is not None else G.img_channels
        self.crop_size, self.preprocess = crop_size, preprocess
        self.fused = fused
        self.loss_log = []

    def broadcasting(self, w):
        return w.repeat([1, self.num_ws, 1])                     # ULA:493-494

    def z_to_w(self, z):
        return self.G.mapping(z, None, truncation_psi=self.truncation_psi)[:, :1, :]   # ULA:459-464

    def forward(self, w):
        if w.ndim == 2:
            w = self.z_to_w(w)
        w_opt = w.detach().clone().to(torch.float32).requires_grad_(True)     # ULA:212
        optim = torch.optim.Adam([w_opt], betas=(0.9, 0.999), lr=self.opt_lr)  # ULA:213
        crop_pos = get_crop_params(self.res, self.crop_size, self.preprocess)  # ULA:216-217 (one window per call)
        self.loss_log = []
        for _ in range(self.num_epochs):                                       # ULA:220
            ws = self.broadcasting(w_opt)
            x = self.G.synthesis(ws, noise_mode='const', fused=self.fused)
            l_lat = calc_loss_latent(ws, self.W, self.w_latent) if self.w_latent > 0 else 0.0
            l_pix = 0.0
            if self.w_pix > 0:
                l_pix = calc_loss_pix(center_crop(x, self.res), center_crop(self.X, self.res),
                                      self.w_pix, self.n_modalities)
            l_disc = 0.0
            if self.w_disc > 0:                                                # ULA:363-371
                l_disc = torch.nn.functional.softplus(-self.D(x, c=None)).mean() * self.w_disc
            l_lpips = 0.0
            if self.w_lpips > 0:                                               # ULA:258-266,387-424
                from . import lpips as olp
                xc = olp.crop(center_crop(x, self.res) if self.preprocess == 'center_random_crop' else x, crop_pos, self.crop_size)
                l_lpips = olp.calc_loss_lpips(self.lpips['state'], xc, self.lpips['bank_feats'], self.w_lpips, self.lpips['taps'],
                                              self.lpips['script'])
            loss = -l_lat - l_pix - l_lpips + l_disc                           # ULA:270
            self.loss_log.append(tuple(float(torch.as_tensor(v).detach()) for v in (l_lat, l_pix, loss, l_disc, l_lpips)))
            optim.zero_grad()
            loss.backward()
            optim.step()
        w_fin = w_opt.detach()
        if self.soft_aug:                                                      # ULA:438-446
            w_aug = self.broadcasting(self.alpha * w_fin + (1 - self.alpha) * w)
        else:                                                                  # ULA:448-454
            w_aug = self.broadcasting(w_fin)
        with torch.no_grad():
            img = self.G.synthesis(w_aug, fused=self.fused)                    # ULA:486-489 (default noise mode)
        return img, w_aug

    def forward_ganrand(self, z):
        """ULA:202-205."""
        w_aug = self.G.mapping(z, None, truncation_psi=self.truncation_psi)
        with torch.no_grad():
            return self.G.synthesis(w_aug, fused=self.fused), w_aug
