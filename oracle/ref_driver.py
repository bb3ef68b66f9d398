"""Drives the REFERENCE's own code (its ``LatentAug.forward`` loop, criteria and ``torch_utils.ops``) on the
synthetic workloads.  TEST / MEASUREMENT INFRASTRUCTURE (see oracle/__init__.py): used by
``oracle/make_golden*.py`` (build container, reads ``/root/reference``) and by bench.py's ``cpu_baseline`` /
``gpu_reference`` / ``--impl reference`` legs (GPU box, reads the git-ignored copy ``baseline/_ref`` that
``oracle/install_ref.py`` makes).  Nothing under ``latentaugment_b200/`` imports this.

What runs from the reference: ``augments/utils/util_latent_aug.py`` ``LatentAug.forward`` (:207-310) with its criteria
(:315-433) and ``torch.optim.Adam``; ``models/stylegan3/torch_utils/ops/{bias_act,upfirdn2d,conv2d_resample,
conv2d_gradfix,fma}.py`` -- on CUDA with the reference's JIT-built plugins (``custom_ops.py:59-155``) + cuDNN, on the
CPU with their ``ref`` paths.  What does NOT come from the reference: the generator *class* (absent from its tree,
SURVEY.md F1) -- ``RefOpsGenerator`` below composes the reference ops per the published StyleGAN2 architecture over
the oracle generator's parameters.
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = '/root/reference'
REF_COPY = os.path.join(ROOT, 'baseline', '_ref')


def find_reference():
    """Path of a usable reference tree: the read-only mount in the build container, else the shipped copy."""
    for root in (REF_SRC, REF_COPY):
        if os.path.exists(os.path.join(root, 'augments', 'utils', 'util_latent_aug.py')):
            return root
    return None


_cached = {}


def import_reference(root=None):
    """Imports the reference modules the hot path needs (third-party plotting / augmentation packages that are not
    installed here are stubbed; none of them is touched by ``LatentAug.forward``)."""
    root = root or find_reference()
    if root is None:
        raise FileNotFoundError('no reference tree: neither /root/reference nor baseline/_ref exists (python -m oracle.install_ref)')
    if root in _cached:
        return _cached[root]
    sys.path[:0] = [root, os.path.join(root, 'models/stylegan3')]
    for name in ['openpyxl', 'matplotlib', 'matplotlib.pyplot', 'albumentations', 'albumentations.pytorch',
                 'kornia', 'kornia.augmentation', 'cv2']:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['albumentations.pytorch'].ToTensorV2 = object
    from torch_utils import custom_ops
    from torch_utils.ops import bias_act, upfirdn2d, conv2d_resample, fma, filtered_lrelu
    from augments.utils import util_latent_aug, util_dataset
    ns = types.SimpleNamespace(root=root, bias_act=bias_act, upfirdn2d=upfirdn2d, conv2d_resample=conv2d_resample, fma=fma,
                               filtered_lrelu=filtered_lrelu, custom_ops=custom_ops, ula=util_latent_aug, uds=util_dataset)
    _cached[root] = ns
    return ns


def init_cuda_plugins(ref):
    """JIT-builds the reference's CUDA plugins the SG2 path uses (bias_act, upfirdn2d).  If a build fails (the sources
    target torch 1.9) the ops are switched to their own ``ref`` implementations -- still the reference's torch path,
    cuDNN convolutions included.  Returns a description string."""
    import glob

    import torch.utils.cpp_extension as ce
    ref.custom_ops.verbosity = 'none'
    status = []
    for mod in (ref.bias_act, ref.upfirdn2d):
        name = mod.__name__.rsplit('.', 1)[-1]
        try:
            mod._init()
            status.append(f'{name}: jit cuda plugin')
            continue
        except ModuleNotFoundError:
            # custom_ops.get_plugin builds with torch.utils.cpp_extension.load(...) and then importlib.import_module()s
            # the plugin by NAME (custom_ops.py:136-139); torch >= 2 no longer leaves the build directory on sys.path, so
            # the import of the freshly built .so fails.  Put the directory it was built in on sys.path and ask again
            # (the second call finds the cached build).  The reference files themselves stay untouched.
            plug = f'{name}_plugin'
            hits = glob.glob(os.path.join(ce._get_build_directory(plug, verbose=False), '*', plug + '*.so'))
            try:
                if not hits:
                    raise
                sys.path.insert(0, os.path.dirname(sorted(hits, key=os.path.getmtime)[-1]))
                mod._init()
                status.append(f'{name}: jit cuda plugin')
                continue
            except Exception as exc:   # noqa: BLE001
                err = exc
        except Exception as exc:       # noqa: BLE001  (any build / load failure)
            err = exc
        mod._init = lambda: False
        status.append(f'{name}: ref ops (plugin build failed: {type(err).__name__})')
    return '; '.join(status)


class RefOpsGenerator(torch.nn.Module):
    """Oracle generator parameters, reference ops.  Fused (grouped-conv) modulation,
    i.e. the eval-mode form the unpickled upstream network runs."""

    def __init__(self, G, ref):
        super().__init__()
        self.G, self.ref = G, ref
        self.z_dim, self.w_dim, self.num_ws = G.z_dim, G.w_dim, G.num_ws
        self.mapping = G.mapping
        self.synthesis = self._synthesis

    def _fc(self, fc, x):
        return torch.addmm((fc.bias * fc.b_gain).unsqueeze(0), x, (fc.weight * fc.w_gain).t())

    def _modconv(self, x, weight, styles, noise, up, padding, f, demod, flip_weight):
        B = x.shape[0]
        O, I, kh, kw = weight.shape
        w = weight.unsqueeze(0) * styles.reshape(B, 1, -1, 1, 1)
        if demod:
            d = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
            w = w * d.reshape(B, -1, 1, 1, 1)
        x = x.reshape(1, -1, *x.shape[2:])
        w = w.reshape(-1, I, kh, kw)
        x = self.ref.conv2d_resample.conv2d_resample(x=x, w=w, f=f, up=up, padding=padding, groups=B,
                                                     flip_weight=flip_weight)
        x = x.reshape(B, -1, *x.shape[2:])
        if noise is not None:
            x = x.add_(noise)
        return x

    def _layer(self, L, x, w, noise_mode):
        styles = self._fc(L.affine, w)
        noise = None
        if noise_mode == 'random':
            noise = torch.randn([x.shape[0], 1, L.res, L.res], device=x.device) * L.noise_strength
        if noise_mode == 'const':
            noise = L.noise_const * L.noise_strength
        x = self._modconv(x, L.weight, styles, noise, L.up, 1, L.resample_filter, True, L.up == 1)
        return self.ref.bias_act.bias_act(x, L.bias, act='lrelu', gain=2 ** 0.5, clamp=L.conv_clamp)

    def _torgb(self, T, x, w):
        styles = self._fc(T.affine, w) * T.w_gain
        x = self._modconv(x, T.weight, styles, None, 1, 0, None, False, True)
        return self.ref.bias_act.bias_act(x, T.bias, clamp=T.conv_clamp)

    def _synthesis(self, ws, noise_mode='random', **_):
        S = self.G.synthesis
        x = img = None
        idx = 0
        for r in S.block_resolutions:
            blk = getattr(S, f'b{r}')
            wi = iter(ws.narrow(1, idx, blk.num_conv + blk.num_torgb).unbind(dim=1))
            idx += blk.num_conv
            if blk.cin == 0:
                x = blk.const.unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
            else:
                x = self._layer(blk.conv0, x, next(wi), noise_mode)
            x = self._layer(blk.conv1, x, next(wi), noise_mode)
            if img is not None:
                img = self.ref.upfirdn2d.upsample2d(img, blk.resample_filter)
            y = self._torgb(blk.torgb, x, next(wi))
            img = img.add_(y) if img is not None else y
        return img


def make_reference_latentaug(ref, Gref, W, X, cfg, steps, w_latent, w_pix, soft_aug=False, alpha=1.0, w_lpips=0.0, w_disc=0.0):
    """The reference's ``LatentAug`` built via ``__new__`` (its ``__init__`` needs zips / pickles that do not exist
    offline) with exactly the attributes its ``forward`` reads."""
    LA = ref.ula.LatentAug
    m = LA.__new__(LA)
    torch.nn.Module.__init__(m)
    m.G = Gref
    m.num_ws, m.w_dim, m.z_dim = Gref.num_ws, Gref.w_dim, Gref.z_dim
    m.batch_size, m.world_size = cfg['batch'], 1
    m.res = cfg['img_resolution']
    m.modalities = [f'm{i}' for i in range(cfg['img_channels'])]
    m.num_epochs, m.opt_lr = steps, 0.01
    m.w_latent, m.w_pix, m.w_lpips, m.w_disc = w_latent, w_pix, w_lpips, w_disc
    m.crop_size, m.preprocess = 64, 'center_random_crop'
    m.soft_aug, m.alpha = soft_aug, alpha
    m.truncation_psi = 1.0
    m.verbose_flag, m.verbose_log = False, False
    m.lpips_script = 'lpips_script'
    m.register_buffer('W', W)
    if X is not None:
        m.register_buffer('X', X)
    return m


def reference_loop(cfg, *, batch, steps, device='cpu', noise_strength=0.0, w_latent=1.0, w_pix=1.0, ref=None,
                   state=None, W=None, X=None, w0=None):
    """(callable running one reference ``LatentAug.forward`` on the seeded synthetic workload, workload dict).
    The callable returns ``(img, w_aug)`` as the reference does.  ``state`` / ``W`` / ``X`` / ``w0`` replace the
    generator parameters, banks and initial codes of the synthetic workload (bench.py passes the product's own, so
    both arms optimise exactly the same problem)."""
    from . import sg2, synthetic
    ref = ref or import_reference()
    if state is None:
        wl = synthetic.make_workload(cfg, noise_strength=noise_strength, batch=batch)
    else:
        c = dict(synthetic.CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg)
        G = sg2.Generator(**synthetic.generator_kwargs(c)).eval().requires_grad_(False)
        G.load_state_dict({k: v.detach().cpu() for k, v in state.items()})
        wl = dict(G=G, W=W.detach().cpu(), X=None if X is None else X.detach().cpu(), w0=w0.detach().cpu(), cfg=c)
    c = dict(wl['cfg'])
    c['batch'] = batch
    G = wl['G'].to(device)
    Gref = RefOpsGenerator(G, ref)
    la = make_reference_latentaug(ref, Gref, wl['W'].to(device), wl['X'].to(device) if wl['X'] is not None else None, c, steps,
                                  w_latent, w_pix)
    w0d = wl['w0'].to(device)

    def run():
        return la(w0d.clone(), ['synthetic'] * batch)
    return run, wl
