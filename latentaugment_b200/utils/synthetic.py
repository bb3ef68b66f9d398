"""Random-init StyleGAN2 generator parameters and synthetic banks for the
throughput runs and for users without a trained pickle.  Product-side: does NOT use oracle/.

Parameter names and layouts follow the reference contract ``models/stylegan3/legacy.py:171-203``:
``mapping.fc{i}.{weight,bias}``, ``mapping.w_avg``, ``synthesis.b{res}.{const,resample_filter}``,
``synthesis.b{res}.{conv0,conv1}.{weight,bias,noise_const,noise_strength,affine.weight,affine.bias}``,
``synthesis.b{res}.torgb.{weight,bias,affine.weight,affine.bias}``.
"""
import math

import torch


def channels_for(img_resolution, channel_base=32768, channel_max=512):
    log2 = int(math.log2(img_resolution))
    assert 2 ** log2 == img_resolution and img_resolution >= 8
    return {4 << b: min(channel_base // (4 << b), channel_max) for b in range(log2 - 1)}


def random_generator_state(img_resolution=256, img_channels=3, w_dim=512, z_dim=512, channel_base=32768, channel_max=512,
                           mapping_layers=8, lr_multiplier=0.01, noise_strength=0.0, seed=0, device='cpu'):
    g = torch.Generator(device='cpu').manual_seed(seed)

    def randn(*shape):
        return torch.randn(shape, generator=g).to(device)
    ch = channels_for(img_resolution, channel_base, channel_max)
    sd = {}
    dims = [z_dim] + [w_dim] * mapping_layers
    for i in range(mapping_layers):
        sd[f'mapping.fc{i}.weight'] = randn(dims[i + 1], dims[i]) / lr_multiplier
        sd[f'mapping.fc{i}.bias'] = torch.zeros(dims[i + 1], device=device)
    sd['mapping.w_avg'] = torch.zeros(w_dim, device=device)
    f1 = torch.tensor([1., 3., 3., 1.])
    fir = (torch.outer(f1, f1) / 64.0).to(device)
    for res, c in ch.items():
        p = f'synthesis.b{res}.'
        sd[p + 'resample_filter'] = fir
        layers = []
        if res == 4:
            sd[p + 'const'] = randn(c, 4, 4)
        else:
            layers.append(('conv0', ch[res // 2]))
        layers.append(('conv1', c))
        for name, cin in layers:
            q = p + name + '.'
            sd[q + 'weight'] = randn(c, cin, 3, 3)
            sd[q + 'bias'] = torch.zeros(c, device=device)
            sd[q + 'noise_const'] = randn(res, res)
            sd[q + 'noise_strength'] = torch.tensor(float(noise_strength), device=device)
            sd[q + 'affine.weight'] = randn(cin, w_dim)
            sd[q + 'affine.bias'] = torch.ones(cin, device=device)
            sd[q + 'resample_filter'] = fir
        q = p + 'torgb.'
        sd[q + 'weight'] = randn(img_channels, c, 1, 1)
        sd[q + 'bias'] = torch.zeros(img_channels, device=device)
        sd[q + 'affine.weight'] = randn(c, w_dim)
        sd[q + 'affine.bias'] = torch.ones(c, device=device)
    return sd


def infer_generator_kwargs(state):
    """(img_resolution, img_channels, w_dim, z_dim) from a reference-named state dict."""
    res = max(int(k.split('.')[1][1:]) for k in state if k.startswith('synthesis.b') and k.endswith('torgb.weight'))
    tw = state[f'synthesis.b{res}.torgb.weight']
    aw = state['synthesis.b4.conv1.affine.weight']
    z_dim = state['mapping.fc0.weight'].shape[1] if 'mapping.fc0.weight' in state else aw.shape[1]
    return dict(img_resolution=res, img_channels=tw.shape[0], w_dim=aw.shape[1], z_dim=z_dim)


def random_discriminator_state(img_resolution=256, img_channels=3, channel_base=32768, channel_max=512, seed=5, device='cpu'):
    """Random-init StyleGAN2 'resnet' discriminator parameters (names per models/stylegan3/legacy.py:267-287)."""
    g = torch.Generator(device='cpu').manual_seed(seed)

    def randn(*shape):
        return torch.randn(shape, generator=g).to(device)
    log2 = int(math.log2(img_resolution))
    res_list = [2 ** i for i in range(log2, 2, -1)]
    ch = {r: min(channel_base // r, channel_max) for r in res_list + [4]}
    f1 = torch.tensor([1., 3., 3., 1.])
    fir = (torch.outer(f1, f1) / 64.0).to(device)
    sd = {}
    for r in res_list:
        p = f'b{r}.'
        c, cn = ch[r], ch[r // 2]
        if r == img_resolution:
            sd[p + 'fromrgb.weight'] = randn(c, img_channels, 1, 1)
            sd[p + 'fromrgb.bias'] = torch.zeros(c, device=device)
        sd[p + 'conv0.weight'] = randn(c, c, 3, 3)
        sd[p + 'conv0.bias'] = torch.zeros(c, device=device)
        sd[p + 'conv1.weight'] = randn(cn, c, 3, 3)
        sd[p + 'conv1.bias'] = torch.zeros(cn, device=device)
        sd[p + 'conv1.resample_filter'] = fir
        sd[p + 'skip.weight'] = randn(cn, c, 1, 1)
    c4 = ch[4]
    sd['b4.conv.weight'] = randn(c4, c4 + 1, 3, 3)
    sd['b4.conv.bias'] = torch.zeros(c4, device=device)
    sd['b4.fc.weight'] = randn(c4, c4 * 16)
    sd['b4.fc.bias'] = torch.zeros(c4, device=device)
    sd['b4.out.weight'] = randn(1, c4)
    sd['b4.out.bias'] = torch.zeros(1, device=device)
    return sd


# ---- perceptual term: VGG16 + LPIPS linear layers (reference criteria/lpips/networks.py:22-32,87-97)
VGG_CFG = [64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512]
VGG_CONV_IDX = [0, 2, 5, 7, 10, 12, 14, 17, 19, 21, 24, 26, 28]
VGG_TAP_LAYERS = (4, 9, 16, 23, 30)
VGG_TAP_CHANNELS = (64, 128, 256, 512, 512)


def random_vgg_state(seed=7, taps=(16, 23, 30)):
    """Seeded random VGG16 / LPIPS-lin parameters with torchvision's names (``features.{i}.weight/bias``) and
    ``lin.{k}.weight`` [1, C, 1, 1] for the used ``taps`` in order -- for throughput runs and users without the
    pretrained files (they cannot be downloaded offline).  He-scaled so activations stay O(1); positive lin weights."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    cin, ci = 3, 0
    for v in VGG_CFG:
        if v == 'M':
            continue
        i = VGG_CONV_IDX[ci]
        sd[f'features.{i}.weight'] = torch.randn([v, cin, 3, 3], generator=g) * math.sqrt(2.0 / (9 * cin))
        sd[f'features.{i}.bias'] = torch.randn([v], generator=g) * 0.05
        cin = v
        ci += 1
    for k, t in enumerate(taps):
        c = VGG_TAP_CHANNELS[VGG_TAP_LAYERS.index(t)]
        sd[f'lin.{k}.weight'] = torch.rand([1, c, 1, 1], generator=g) * 0.5 + 0.05
    return sd
