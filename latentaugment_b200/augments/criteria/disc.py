"""Realism term (reference ``calc_loss_disc``, util_latent_aug.py:363-371): ``softplus(-D(x, c=None)).mean() * w_disc``
with the StyleGAN2 'resnet' discriminator; enters the objective with a PLUS sign (:270).  Inside the captured loop the
term runs fused (csrc/disc.cu); this plugin attaches the discriminator to an engine and evaluates the same
quantity stand-alone."""
import torch


class DiscriminatorCriterion:
    name, sign = 'disc', +1.0

    def __init__(self, weight=1.0):
        self.weight = float(weight)
        self.engine = None

    def attach(self, engine, state, conv_clamp=256.0, mbstd_group_size=4):
        """``state``: reference-named discriminator parameters (models/stylegan3/legacy.py:267-287)."""
        engine.set_discriminator(state, conv_clamp=conv_clamp, mbstd_group_size=mbstd_group_size)
        self.engine = engine

    def forward(self, x, engine=None):
        """x [batch, C, res, res] -> scalar loss (weight applied)."""
        e = engine or self.engine
        logits = e.disc_logits(x)
        return torch.nn.functional.softplus(-logits).mean() * self.weight

    __call__ = forward
