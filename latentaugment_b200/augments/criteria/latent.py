"""Latent-diversity term (reference ``calc_loss_latent``, util_latent_aug.py:427-433):
``w_latent * mean_{i,j} |ws_i - W_j|^2 / (num_ws * w_dim)``; enters the objective with a minus sign."""
import math

from ...engine import pairwise_sqdist


class LatentCriterion:
    name, sign = 'latent', -1.0

    def __init__(self, weight=1.0):
        self.weight = float(weight)

    def attach(self, engine, bank):
        engine.set_latent_bank(bank)

    def forward(self, ws, bank):
        """Stand-alone value on the GPU: l2_loss_vectorized(ws, W) * w_latent (:430)."""
        D = pairwise_sqdist(ws, bank)                       # [bank, batch]
        return D.sum() / (D.shape[0] * D.shape[1]) / math.prod(bank.shape[1:]) * self.weight

    __call__ = forward
