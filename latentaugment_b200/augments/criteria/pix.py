"""Pixel term (reference ``calc_loss_pix``, util_latent_aug.py:373-385): per modality, mean over
all (sample, bank image) pairs of the squared L2 between centre crops / (h*w), * w_pix, averaged
over modalities; enters the objective with a minus sign."""
import math

from ...engine import pairwise_sqdist


def center_crop_bounds(res):
    """torchvision CenterCrop(int(sqrt(res^2/2))) as used by util_dataset.py:317-323 -> (offset, size)."""
    size = int(math.sqrt((res * res) / 2))
    return int(round((res - size) / 2.0)), size


class PixelCriterion:
    name, sign = 'pix', -1.0

    def __init__(self, weight=1.0):
        self.weight = float(weight)

    def attach(self, engine, bank):
        engine.set_image_bank(bank)

    def forward(self, x, bank):
        """x [n,C,res,res], bank [m,C,res,res] (uncropped): crops both, like :253."""
        off, size = center_crop_bounds(x.shape[-1])
        xc = x[:, :, off:off + size, off:off + size]
        bc = bank[:, :, off:off + size, off:off + size]
        C = x.shape[1]
        loss = 0.0
        for c in range(C):
            D = pairwise_sqdist(xc[:, c].contiguous(), bc[:, c].contiguous())
            loss = loss + D.sum() / (D.shape[0] * D.shape[1]) / (size * size) * self.weight
        return loss / C

    __call__ = forward
