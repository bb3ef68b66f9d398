"""Perceptual term (reference ``calc_loss_lpips_torchscript`` / ``calc_loss_lpips_tr``, util_latent_aug.py:387-424;
``criteria/lpips/{lpips,networks,utils}.py``): per modality, a ``crop_size_aug`` window of the synthetic image (one
position per forward call, inside the centre crop) -> 3 channels -> VGG16 -> channel-normalised activations at the
taps -> LPIPS distance to every bank image; ``* w_lpips``, averaged over modalities; enters the objective with a minus sign.

``--lpips_script lpips_script`` (default) is the NVIDIA TorchScript form: all five taps, loss = mean over (sample, bank)
pairs of the squared feature distance (:400-405).  Any other value is the in-tree ``LPIPS`` form: taps 16/23/30
(networks.py:94), loss = sum over pairs / bank size (``forward_tr``, lpips.py:58-68, with its broadcast fixed to pairs).
The pretrained weights (NVIDIA ``vgg16.pt`` / torchvision VGG16 + the LPIPS lin layers) are not downloadable here:
``--vgg_state FILE`` loads them (torchvision ``features.*`` names + ``lin.{k}.weight``), ``--synthetic`` draws seeded
random ones.
"""
TAPS_SCRIPT = (4, 9, 16, 23, 30)
TAPS_INTREE = (16, 23, 30)


def taps_and_norm(lpips_script):
    """(tap layers, pair-normaliser mode of the C ABI) for the reference's ``--lpips_script`` flag."""
    return (TAPS_SCRIPT, 0) if lpips_script == 'lpips_script' else (TAPS_INTREE, 1)


class PerceptualCriterion:
    name, sign = 'lpips', -1.0

    def __init__(self, weight=1.0):
        self.weight = float(weight)
        self.engine = None
        self.norm_mode = 0

    def attach(self, engine, vgg_state, bank_crops, lpips_script='lpips_script', crop_size=64):
        taps, self.norm_mode = taps_and_norm(lpips_script)
        engine.set_lpips(vgg_state, taps=taps, crop_size=crop_size)
        engine.set_feature_bank(bank_crops)
        self.engine = engine

    def forward(self, x, crop_pos):
        """Stand-alone value on the GPU for one engine's batch shard ``x`` [n, C, res, res] (uncropped) and the crop
        position ``(x, y)`` of this call (util_dataset.get_params)."""
        return self.engine.lpips_loss_grad(x, crop_pos, self.weight, self.norm_mode)[0]

    __call__ = forward
