"""Criteria (loss-term) plugins of the latent-optimisation loop.

The reference has no formal registry: terms are methods ``calc_loss_{latent,pix,lpips_*,disc}``
of ``LatentAug`` gated by ``w_* > 0`` (augments/utils/util_latent_aug.py:232-267) with modules
under ``augments/criteria/<name>/``.  Here each term is a plugin class with the weight and sign
handled by the loop exactly as ``loss = -latent - pix - lpips + disc`` (:270).

Inside the captured CUDA loop the latent and pixel terms run fused (bank-moment form,
csrc/kernels.cu) and so do the discriminator term (csrc/disc.cu) and the perceptual term (csrc/lpips.cu); the plugin objects configure the engine and evaluate the same quantity
stand-alone through the pairwise-distance kernel (the reference's ``l2_loss_vectorized``).
"""
from .disc import DiscriminatorCriterion
from .latent import LatentCriterion
from .lpips import PerceptualCriterion
from .pix import PixelCriterion

REGISTRY = {'latent': LatentCriterion, 'pix': PixelCriterion, 'lpips': PerceptualCriterion, 'disc': DiscriminatorCriterion}
UNAVAILABLE = {}


def find_criterion_using_name(name):
    if name in REGISTRY:
        return REGISTRY[name]
    if name in UNAVAILABLE:
        raise NotImplementedError(f'criterion {name!r}: {UNAVAILABLE[name]}')
    raise KeyError(name)


def create_criteria(opt):
    """{name: instance} for every term with a positive weight."""
    out = {}
    for name in ('latent', 'pix', 'lpips', 'disc'):
        w = float(getattr(opt, 'w_' + name, 0.0))
        if w > 0:
            out[name] = find_criterion_using_name(name)(w)
    return out
