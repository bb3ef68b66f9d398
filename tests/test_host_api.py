"""CPU tests of the host side: option parsing, plugin registry, criteria registry, the C-ABI
library (loads; exports every symbol include/*.h declares -- no compute without a GPU), and the
multi-process plumbing under gloo with world_size 2."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _opts(extra=(), args=None):
    from latentaugment_b200.options.aug_options import AugOptions
    argv = ['--aug', 'latent', '--synthetic', '--batch_size', '4', '--no_log'] + list(extra)
    return AugOptions().parse(args=args, argv=argv)


def test_options_defaults_match_reference():
    """Flag names / defaults of reference augments/latent_aug.py:58-96 and options/base_options.py:28-39."""
    opt = _opts()
    ref = dict(gpu_ids_aug='0', img_resolution=256, truncation_psi=1.0, rand_aug=False, lower_bound_clip=False, step_img=20,
               step_w=5, lpips_script='lpips_script', opt_num_epochs=10, opt_lr=0.01, init_w='random', crop_size_aug=64,
               preprocess_aug='center_random_crop', w_pix=1.0, w_lpips=1.0, w_latent=1.0, w_disc=1.0, p_thres=1.0,
               soft_aug=False, alpha=1.0, verbose_log=False, phase='train', load_size=256, dataset_mode='pelvis2.1')
    for k, v in ref.items():
        assert getattr(opt, k) == v, k
    assert opt.gpu_ids == [0] and opt.isTrain is True
    assert opt.name.startswith('experiment_name-n_imgs_0-opt_lr_0.01-opt_num_epochs_10-w_latent_1.0')


def test_options_dict_overrides_like_reference():
    opt = _opts(args={'p_thres': 0.0, 'opt_num_epochs': 6, 'opt_lr': 0.1, 'w_lpips': 0.0, 'w_disc': 0.0, 'init_w': 'inv', 'n_imgs': 7})
    assert (opt.p_thres, opt.opt_num_epochs, opt.opt_lr, opt.w_lpips, opt.w_disc, opt.init_w, opt.n_imgs) == (0.0, 6, 0.1, 0.0, 0.0, 'inv', 7)
    opt = _opts(['--rand_aug'], args={'truncation_psi': 0.5, 'opt_num_epochs': 3})
    assert opt.truncation_psi == 0.5 and opt.opt_num_epochs == 10      # rand_aug ignores the loop overrides (base_options.py:118-122)
    assert 'truncation_psi_0.5' in opt.name


def test_plugin_registry():
    from latentaugment_b200 import augments
    from latentaugment_b200.augments.base_aug import BaseAugment
    cls = augments.find_augment_using_name('latent')
    assert cls.__name__ == 'LatentAugment' and issubclass(cls, BaseAugment)
    assert augments.get_option_setter('latent') is cls.modify_commandline_options
    with pytest.raises(ImportError):
        augments.find_augment_using_name('geometric')          # out of scope


def test_criteria_registry():
    from latentaugment_b200.augments import criteria
    opt = _opts(args={'w_lpips': 0.0, 'w_disc': 0.0})
    c = criteria.create_criteria(opt)
    assert sorted(c) == ['latent', 'pix'] and c['latent'].sign == -1.0 and c['pix'].weight == 1.0
    full = criteria.create_criteria(_opts())                   # reference defaults: all four weights are 1
    assert sorted(full) == ['disc', 'latent', 'lpips', 'pix'] and full['lpips'].sign == -1.0 and full['disc'].sign == 1.0


def test_no_cpu_fallback():
    from latentaugment_b200 import LatentAugmentError
    from latentaugment_b200.augments import create_augment
    if torch.cuda.is_available():
        pytest.skip('needs a machine without a GPU')
    opt = _opts(args={'w_lpips': 0.0, 'w_disc': 0.0})
    with pytest.raises(LatentAugmentError):
        create_augment(opt)


def test_val_phase_passthrough():
    from latentaugment_b200.augments import create_augment
    opt = _opts(['--phase', 'val'])
    aug = create_augment(opt)
    a, b = torch.rand(4, 1, 8, 8), torch.rand(4, 1, 8, 8)
    aug.set_input({'A': a, 'B': b, 'A_paths': ['x'] * 4, 'B_paths': ['x'] * 4})
    aug.forward()
    out = aug.get_output()
    assert torch.equal(out['A'], a) and torch.equal(out['B'], b) and len(aug.stats_time) == 1


def test_cabi_exports_every_declared_symbol():
    from latentaugment_b200 import _build, _lib
    lib_path = _build.build()
    header = open(os.path.join(ROOT, 'include', 'latentaugment_b200.h')).read()
    header = re.sub(r'/\*.*?\*/', '', header, flags=re.S)
    declared = set(re.findall(r'\b(la_[a-z0-9_]+)\s*\(', header))
    assert {'la_engine_create', 'la_augment', 'la_synthesis', 'la_mapping', 'la_nearest_codes', 'la_pairwise_sqdist'} <= declared
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in the header but not exported'
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().la_version() == _lib.ABI_VERSION == int(re.search(r"LA_ABI_VERSION (\d+)", open(os.path.join(ROOT, "include", "latentaugment_b200.h")).read()).group(1))


def test_synthetic_state_follows_reference_naming():
    from latentaugment_b200.utils import synthetic
    sd = synthetic.random_generator_state(img_resolution=32, img_channels=2, channel_base=2048, channel_max=64)
    assert sd['synthesis.b4.const'].shape == (64, 4, 4)
    assert sd['synthesis.b8.conv0.weight'].shape == (64, 64, 3, 3) and sd['synthesis.b8.conv0.affine.weight'].shape == (64, 512)
    assert bool((sd['synthesis.b16.conv1.affine.bias'] == 1).all())            # legacy.py:183,189,195,199
    assert sd['synthesis.b32.torgb.weight'].shape == (2, 64, 1, 1) and 'synthesis.b4.conv0.weight' not in sd
    assert synthetic.infer_generator_kwargs(sd) == dict(img_resolution=32, img_channels=2, w_dim=512, z_dim=512)
    from oracle import sg2                                                     # same module tree as the oracle generator
    G = sg2.Generator(img_resolution=32, img_channels=2, channel_base=2048, channel_max=64)
    assert set(G.state_dict()) == set(sd)


def _merge_cpu(dist_, idx):
    s, n, k = dist_.shape
    d = dist_.permute(1, 0, 2).reshape(n, s * k)
    i = idx.permute(1, 0, 2).reshape(n, s * k)
    key = torch.argsort(i, dim=1, stable=True)
    d, i = d.gather(1, key), i.gather(1, key)
    order = torch.argsort(d, dim=1, stable=True)[:, :k]
    return d.gather(1, order), i.gather(1, order)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from latentaugment_b200 import parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    gen = torch.Generator().manual_seed(0)
    Y = torch.randn([101, 16], generator=gen)
    X = torch.randn([6, 16], generator=gen)
    b, e = parallel.shard_range(101, rank, world)

    def nearest(Xq, k):          # exact search on this rank's rows, global indices
        D = torch.cdist(Xq, Y[b:e]).square()
        d, i = torch.sort(D, dim=1, stable=True)
        return d[:, :k].contiguous(), (i[:, :k] + b).contiguous()
    xb, xe = parallel.shard_range(6, rank, world)
    Xall = parallel.all_gather_queries(X[xb:xe].contiguous())
    d, i = parallel.sharded_nearest_codes(nearest, Xall, 3, merge_fn=_merge_cpu)
    D = torch.cdist(X, Y).square()
    dr, ir = torch.sort(D, dim=1, stable=True)
    ok = torch.equal(Xall, X) and torch.equal(i, ir[:, :3]) and torch.allclose(d, dr[:, :3])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_nearest_codes_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_range_covers_everything():
    from latentaugment_b200.parallel import shard_range
    for n in (1, 7, 128, 1000003):
        for world in (1, 2, 3, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n and all(r[k][1] == r[k + 1][0] for k in range(world - 1))


def test_feature_bank_crops_follow_reference_draw_order():
    """One fresh get_params draw (two random.randint calls) per (modality, bank image), modality-major -- the order in
    which the reference builds fea_{mode} (util_latent_aug.py:160-169,564-579)."""
    import random

    import torch

    from latentaugment_b200.augments.criteria.pix import center_crop_bounds
    from latentaugment_b200.augments.utils.util_latent_aug import feature_bank_crops
    X = torch.arange(3 * 2 * 128 * 128, dtype=torch.float32).reshape(3, 2, 128, 128)
    random.seed(4)
    crops = feature_bank_crops(X, 128, 64)
    random.seed(4)
    off, size = center_crop_bounds(128)
    for c in range(2):
        for m in range(3):
            x = random.randint(0, size - 64)
            y = random.randint(0, size - 64)
            assert torch.equal(crops[m, c], X[m, c, off + y:off + y + 64, off + x:off + x + 64])


def test_lpips_flag_selects_taps_and_normaliser():
    from latentaugment_b200.augments.criteria import REGISTRY, create_criteria
    from latentaugment_b200.augments.criteria.lpips import taps_and_norm
    assert taps_and_norm('lpips_script') == ((4, 9, 16, 23, 30), 0)
    assert taps_and_norm('lpips') == ((16, 23, 30), 1)
    import types
    crit = create_criteria(types.SimpleNamespace(w_latent=1.0, w_pix=0.0, w_lpips=1.0, w_disc=0.0))
    assert set(crit) == {'latent', 'lpips'} and isinstance(crit['lpips'], REGISTRY['lpips'])


def test_lookahead_loop_passthrough_on_cpu():
    """iterate() = the reference's caller loop with one batch of look-ahead; in the val phase (no augmentation, no GPU)
    it must hand back every batch unchanged and in order, like set_input / forward / get_output."""
    from latentaugment_b200.augments import create_augment
    aug = create_augment(_opts(['--phase', 'val']))
    batches = [{'A': torch.full((4, 1, 8, 8), float(i)), 'B': torch.full((4, 1, 8, 8), -float(i)), 'A_paths': [f'p{i}'] * 4, 'B_paths': [f'p{i}'] * 4}
               for i in range(3)]
    got = list(aug.iterate(iter(batches)))
    assert len(got) == 3 and len(aug.stats_time) == 3
    for (data, out), ref in zip(got, batches):
        assert data is ref and out['A_paths'] == ref['A_paths']
        assert torch.equal(out['A'], ref['A']) and torch.equal(out['B'], ref['B'])
    assert list(aug.iterate(iter([]))) == []


class _CpuBank:
    """LatentBank stand-in on CPU: exact search of this rank's rows, GLOBAL indices, writes into ``out`` like the real one."""

    def __init__(self, Y, index_offset=0):
        self.Y, self.K, self.off = Y, Y.shape[1], index_offset

    def nearest(self, X, k=1, out=None):
        d, i = torch.sort(torch.cdist(X, self.Y).square(), dim=1, stable=True)
        d, i = d[:, :k].contiguous(), (i[:, :k] + self.off).contiguous()
        if out is not None:
            out[0].copy_(d)
            out[1].copy_(i)
            return out
        return d, i


class _CpuLib:
    """The one C-ABI call of ShardedNearest._exchange, executed on host memory: reads the gathered records through the raw
    pointers / strides it is given (so the record layout and the pointer arithmetic of the caller are what is tested)."""

    @staticmethod
    def la_merge_topk_strided(d_dist, d_idx, shards, n, k, dist_stride, idx_stride, d_out_dist, d_out_idx, stream):
        import numpy as np
        f32 = ctypes.POINTER(ctypes.c_float)
        i64 = ctypes.POINTER(ctypes.c_longlong)
        dist = np.ctypeslib.as_array(ctypes.cast(d_dist, f32), shape=((shards - 1) * dist_stride + n * k,))
        idx = np.ctypeslib.as_array(ctypes.cast(d_idx, i64), shape=((shards - 1) * idx_stride + n * k,))
        od = np.ctypeslib.as_array(ctypes.cast(d_out_dist, f32), shape=(n * k,))
        oi = np.ctypeslib.as_array(ctypes.cast(d_out_idx, i64), shape=(n * k,))
        for q in range(n):
            cand = sorted((float(dist[s * dist_stride + q * k + t]), int(idx[s * idx_stride + q * k + t])) for s in range(shards) for t in range(k))
            for t in range(k):
                od[q * k + t], oi[q * k + t] = cand[t]
        return 0


def _sharded_worker(rank, world, port, q):
    import contextlib

    import torch.distributed as dist

    from latentaugment_b200 import _lib, engine, parallel
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    # CUDA-only pieces of the class, stubbed: the library call, stream / device plumbing (use_graph=False: no capture)
    _lib.load = lambda: _CpuLib
    _lib.check = lambda rc: None
    engine._stream_ptr = lambda dev: None
    engine._ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.device = lambda dev: contextlib.nullcontext()
    gen = torch.Generator().manual_seed(0)
    Y = torch.randn([96, 16], generator=gen)
    X = torch.randn([7, 16], generator=gen)
    b, e = parallel.shard_range(96, rank, world)
    s = parallel.ShardedNearest(_CpuBank(Y[b:e], b), 7, 3, use_graph=False)
    d, i = s(X)
    d2, i2 = s(X.clone())                        # static buffers: a second call gives the same answer
    dr, ir = torch.sort(torch.cdist(X, Y).square(), dim=1, stable=True)
    ok = (s.world == world and s.rec_bytes == 7 * 3 * 16 and torch.equal(i, ir[:, :3]) and torch.allclose(d, dr[:, :3])
          and torch.equal(i2, ir[:, :3]) and i.dtype == torch.int64 and d.dtype == torch.float32)
    s.close()
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_nearest_class_gloo_world2():
    """parallel.ShardedNearest under a 2-rank gloo group: local search into the packed (dist | pad | idx) record, ONE
    all_gather_into_tensor of it, the merge call with the strides / offsets of that record."""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _generator_desc(res, img_c, cb=32768, cm=512):
    import math

    from latentaugment_b200 import _lib
    from latentaugment_b200.utils.synthetic import channels_for
    d = _lib.GeneratorDesc()
    d.img_resolution, d.img_channels, d.w_dim, d.z_dim = res, img_c, 512, 512
    d.num_blocks = int(math.log2(res)) - 1
    d.conv_clamp = 256.0
    ch = channels_for(res, cb, cm)
    for b in range(d.num_blocks):
        d.channels[b] = ch[4 << b]
    d.mapping_layers, d.mapping_lr_multiplier = 8, 0.01
    return d


def test_workspace_planners_run_on_the_host_for_the_benchmark_shapes():
    """The ``*_workspace_bytes`` entry points are pure host planning (no device needed): they accept the shapes of
    BASELINE.json's configs, scale with the batch, and reject shapes the kernels do not cover."""
    from latentaugment_b200 import _lib
    lib = _lib.load()
    GiB = 2.0 ** 30

    def engine_bytes(d, batch, precision):
        n = ctypes.c_size_t(0)
        rc = lib.la_engine_workspace_bytes(ctypes.byref(d), batch, _lib.PRECISION[precision], ctypes.byref(n))
        return rc, n.value
    c2 = _generator_desc(256, 3)
    rc, b32 = engine_bytes(c2, 32, 'bf16')
    assert rc == 0 and 3.5 * GiB < b32 < 6 * GiB                       # 4.7 GiB
    rc, p32 = engine_bytes(c2, 32, 'fp32_parity')
    assert rc == 0 and 1.7 * b32 < p32 < 2.2 * b32                     # hi + lo planes of every activation-like tensor
    rc, b16 = engine_bytes(c2, 16, 'bf16')
    assert rc == 0 and 0.45 * b32 < b16 < 0.6 * b32
    c3 = _generator_desc(512, 3)
    rc, big = engine_bytes(c3, 128, 'bf16')
    assert rc == 0 and 30 * GiB < big < 45 * GiB                       # fits one B200 (180 GB) with room for D and VGG
    rc, _ = engine_bytes(_generator_desc(128, 1), 4, 'fp32_parity')
    assert rc == 0
    bad = _generator_desc(256, 3)
    bad.channels[3] = 48                                               # channel counts must be multiples of 64
    assert engine_bytes(bad, 32, 'bf16')[0] != 0 and lib.la_last_error()
    assert engine_bytes(c2, 0, 'bf16')[0] != 0
    # nearest-code search: candidate lists (2 per 32 codes per padded query) dominate
    n = ctypes.c_size_t(0)
    assert lib.la_nearest_codes_workspace_bytes(1024, 131072, 512, 4, ctypes.byref(n)) == 0
    cand = 1024 * (131072 // 32) * 2 * 8
    assert cand <= n.value < cand + 8 * 2 ** 20
    assert lib.la_nearest_codes_workspace_bytes(1024, 131072, 512, 9, ctypes.byref(n)) != 0        # k <= 8
    # perceptual term: workspace grows with batch x modalities
    v = _lib.VggDesc()
    v.crop_size = 64
    dummy = ctypes.create_string_buffer(64)                            # planning reads which pointers are set, never what they point to
    for i in range(_lib.LA_VGG_CONVS):
        v.d_conv_weight[i] = v.d_conv_bias[i] = ctypes.addressof(dummy)
    for k in (2, 3, 4):                                                # the in-tree taps relu3_3, relu4_3, relu5_3
        v.d_lin_weight[k] = ctypes.addressof(dummy)
    a, b = ctypes.c_size_t(0), ctypes.c_size_t(0)
    assert lib.la_lpips_workspace_bytes(ctypes.byref(v), 32, 3, _lib.PRECISION['bf16'], ctypes.byref(a)) == 0
    assert lib.la_lpips_workspace_bytes(ctypes.byref(v), 32, 1, _lib.PRECISION['bf16'], ctypes.byref(b)) == 0
    assert a.value > b.value > 0
