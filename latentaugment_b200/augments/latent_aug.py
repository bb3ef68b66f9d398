"""Drop-in boundary class: B200-native counterpart of the reference ``augments/latent_aug.py``
(``LatentAugment``: options :45-98, constructor :100-157, ``set_input`` :171-180,
``get_output`` :182-203, ``get_latent_output/input`` :205-235, ``forward`` :237-276,
``sample_from_randn / sample_from_inversion`` :306-324)."""
import random
import time

import torch

from .base_aug import BaseAugment
from .utils import util_latent_aug


def reverse_broadcasting(latent):
    return latent[:, :1, :]


def set_gpu_ids(gpu_ids):
    out = []
    for s in str(gpu_ids).split(','):
        i = int(s)
        if i >= 0:
            out.append(i)
    return out


def _str2bool(v):
    return str(v).lower() in ('1', 'true', 'yes', 'y')


class LatentAugment(BaseAugment):
    @staticmethod
    def modify_commandline_options(parser, is_train):
        # every flag of the reference (latent_aug.py:58-96), same names / types / defaults; the three
        # directory flags are optional here because synthetic mode needs none of them
        parser.add_argument('--model_dir', metavar='DIR', default='')
        parser.add_argument('--interim_dir', metavar='DIR', default='')
        parser.add_argument('--gpu_ids_aug', type=str, default='0')
        parser.add_argument('--dataset_aug', default='Pelvis_2.1_repo_no_mask')
        parser.add_argument('--dataset_name_aug', default='Pelvis_2.1_repo_no_mask-num-375_train-0.70_val-0.20_test-0.10')
        parser.add_argument('--modalities_aug', default='MR_nonrigid_CT,MR_MR_T2')
        parser.add_argument('--img_resolution', type=int, default=256)
        parser.add_argument('--exp_stylegan', default='00003')
        parser.add_argument('--network_pkl_stylegan', default='network-snapshot-005320.pkl')
        parser.add_argument('--dataset_w_name', default='Pelvis_2.1_repo_no_mask-num-375_train-0.70_val-0.20_test-0.10-expinv_00001')
        parser.add_argument('--exp_inv', default='00001')
        parser.add_argument('--network_pkl_inv', default='')
        parser.add_argument('--truncation_psi', type=float, default=1.0)
        parser.add_argument('--rand_aug', action='store_true')
        parser.add_argument('--lower_bound_clip', action='store_true')
        parser.add_argument('--step_img', type=int, default=20)
        parser.add_argument('--step_w', type=int, default=5)
        parser.add_argument('--lpips_script', type=str, default='lpips_script')
        parser.add_argument('--opt_num_epochs', type=int, default=10)
        parser.add_argument('--opt_lr', type=float, default=0.01)
        parser.add_argument('--init_w', type=str, default='random')
        parser.add_argument('--crop_size_aug', type=int, default=64)
        parser.add_argument('--preprocess_aug', type=str, default='center_random_crop')
        parser.add_argument('--w_pix', type=float, default=1.0)
        parser.add_argument('--w_lpips', type=float, default=1.0)
        parser.add_argument('--w_latent', type=float, default=1.0)
        parser.add_argument('--w_disc', type=float, default=1.0)
        parser.add_argument('--p_thres', type=float, default=1.0)
        parser.add_argument('--soft_aug', type=_str2bool, default=False)
        parser.add_argument('--alpha', type=float, default=1.0)
        parser.add_argument('--verbose_log', type=_str2bool, default=False)
        # ---- additions of this implementation
        parser.add_argument('--precision', type=str, default='fp32_parity', choices=['fp32_parity', 'bf16'],
                            help='tensor-core operand precision (fp32_parity = split-bf16, rel-L2 1e-3; bf16 = 1e-2)')
        parser.add_argument('--micro_batches', type=int, default=1, help='cut every GPU shard into k concurrently running parts (own stream + graph each)')
        parser.add_argument('--generator_state', type=str, default='', help='torch state_dict file with the reference parameter names')
        parser.add_argument('--discriminator_state', type=str, default='', help='torch state_dict file of the StyleGAN2 discriminator (needed when w_disc > 0)')
        parser.add_argument('--vgg_state', type=str, default='', help='torch state_dict file of VGG16 (torchvision features.* names) + LPIPS lin layers lin.{k}.weight (needed when w_lpips > 0)')
        parser.add_argument('--latent_bank', type=str, default='', help='tensor file [M, num_ws, w_dim] or [M, w_dim]')
        parser.add_argument('--image_bank', type=str, default='', help='tensor file [M, C, res, res] in [-1, 1]')
        parser.add_argument('--inverted_codes', type=str, default='', help="file with {'names': [...], 'codes': [N, w_dim]}")
        parser.add_argument('--synthetic', action='store_true', help='random-init generator + synthetic banks (SURVEY.md §8d)')
        parser.add_argument('--synthetic_channels', type=int, default=3)
        parser.add_argument('--synthetic_channel_base', type=int, default=32768)
        parser.add_argument('--synthetic_channel_max', type=int, default=512)
        parser.add_argument('--synthetic_bank', type=int, default=4096)
        parser.add_argument('--synthetic_img_bank', type=int, default=64)
        parser.add_argument('--synthetic_codes', type=int, default=1024)
        return parser

    def __init__(self, opt, **core_kwargs):
        BaseAugment.__init__(self, opt)
        self.gpu_ids_aug = set_gpu_ids(opt.gpu_ids_aug)
        self.device = torch.device('cuda:{}'.format(self.gpu_ids_aug[0])) if self.gpu_ids_aug else torch.device('cpu')
        self.phase = opt.phase
        self.batch_size = opt.batch_size
        self.rand_aug = opt.rand_aug
        self.lower_bound_clip = opt.lower_bound_clip
        self.p_thres = opt.p_thres
        self.init_w = opt.init_w
        self.verbose_log = opt.verbose_log
        self.stats_time = []
        self._host_out, self._pending, self._pending_src = {}, None, None
        self._sync_forward, self._passthrough = True, False
        if self.phase == 'train':
            print('\nTrain phase.')
            if self.rand_aug:                                # :126-137
                print('Random GAN augmentation! Disable latent aug parameters.')
                opt.w_pix = opt.w_lpips = opt.w_latent = opt.w_disc = 0.0
                opt.init_w = 'random'
                self.init_w = opt.init_w
                opt.opt_num_epochs = 0
                opt.soft_aug = False
            if self.lower_bound_clip:
                print('Clip pixel values under -1 to -1.')
            self.latent_aug = util_latent_aug.define_latentaugment(
                module_name='latent_aug', phase=opt.phase, opt=opt, save_dir=self.save_dir, gpu_ids=self.gpu_ids_aug,
                **core_kwargs)
            self.stats_dataset_w = self.latent_aug.module.stats_dataset_w
            self.num_ws = self.latent_aug.module.num_ws
            self.w_dim = self.latent_aug.module.w_dim
            self.z_dim = self.latent_aug.module.z_dim
        elif self.phase in ['val', 'test']:
            print('\nVal/Test phase.\nAll augmentation disabled.')
        else:
            raise NotImplementedError

    def set_input(self, data):
        assert data['A_paths'] == data['B_paths']
        self.real_A = data['A']
        self.real_B = data['B']
        self.fname = data['A_paths']
        self._real_AB = None           # (the reference concatenates here, :176; only the pass-through branch reads it -> built on demand)

    @property
    def real_AB(self):
        if self._real_AB is None:
            self._real_AB = torch.cat((self.real_A, self.real_B), dim=1)
        return self._real_AB

    # ---- output path (reference get_output :182-203 does ``.detach().cpu()``)
    # forward() enqueues the device->host copy of the augmented batch right behind the final synthesis, on a copy
    # stream, into one of TWO pinned buffers per shape (no per-call cudaHostAlloc); get_output() waits for that copy's
    # event only.  With the reference's call order (set_input, forward, get_output) the copy is already in flight
    # when get_output is called; with the look-ahead loop ``iterate()`` the copy of batch t runs while batch t+1
    # computes.  A returned dict stays valid until the next-but-one forward().
    def _start_d2h(self, t, tag='img'):
        """-> (host tensor, event or None); ``tag`` separates the staging slots of tensors with equal shapes."""
        if not t.is_cuda:
            pending = (t.detach(), None)
            if tag == 'img':
                self._pending = pending
            return pending
        key = (tag, tuple(t.shape), t.dtype, t.device)
        slot = self._host_out.get(key)
        if slot is None:
            slot = self._host_out[key] = {'buf': [torch.empty(t.shape, dtype=t.dtype).pin_memory() for _ in range(2)], 'i': 0,
                                          'stream': torch.cuda.Stream(t.device), 'ev': [torch.cuda.Event(), torch.cuda.Event()]}
        slot['i'] ^= 1
        i = slot['i']
        cs = slot['stream']
        cs.wait_stream(torch.cuda.current_stream(t.device))
        with torch.cuda.stream(cs):
            slot['buf'][i].copy_(t.detach(), non_blocking=True)
            slot['ev'][i].record(cs)
        t.record_stream(cs)
        pending = (slot['buf'][i], slot['ev'][i])
        if tag == 'img':
            self._pending = pending
        return pending

    def _host_output(self):
        if self._pending is None or self._pending_src is not self.real_AB_aug:
            self._start_d2h(self.real_AB_aug)
            self._pending_src = self.real_AB_aug
        buf, ev = self._pending
        if ev is not None:
            ev.synchronize()
        return buf

    def get_output(self):
        real_AB_aug = self._host_output()
        real_A_aug = real_AB_aug[:, 0, :, :].unsqueeze(dim=1)
        real_B_aug = real_AB_aug[:, min(1, real_AB_aug.shape[1] - 1), :, :].unsqueeze(dim=1)
        if self.lower_bound_clip:
            real_A_aug = torch.clamp(real_A_aug, min=-1.0, max=None)
            real_B_aug = torch.clamp(real_B_aug, min=-1.0, max=None)
        return {'A': real_A_aug, 'B': real_B_aug, 'A_paths': self.fname, 'B_paths': self.fname}

    def get_latent_output(self):
        w_aug = reverse_broadcasting(self.w_AB_aug).detach().cpu().numpy().squeeze()
        return {'w': w_aug, 'paths': self.fname if not self.rand_aug else ''}

    def get_latent_input(self):
        w = self.w_AB.detach().cpu().numpy().squeeze()
        return {'w': w, 'paths': self.fname if not self.rand_aug else ''}

    def forward(self):
        since = time.time()
        if random.random() > self.p_thres and self.phase == 'train':
            self._passthrough = False
            if self.rand_aug:
                w_AB = self.sample_from_randn().to(self.device)
                self.real_AB_aug, self.w_AB_aug = self.latent_aug.module.forward_ganrand(w_AB)
                self.w_AB = self.w_AB_aug
            else:
                if self.init_w == 'random':
                    raise NotImplementedError            # as the reference (:253-255)
                elif self.init_w == 'inv':
                    self.w_AB = self.sample_from_inversion(self.fname)
                else:
                    raise NotImplementedError
                self.w_AB = self.w_AB.to(self.device, non_blocking=True)
                self.real_AB_aug, self.w_AB_aug = self.latent_aug(self.w_AB, self.fname)
            self._start_d2h(self.real_AB_aug)
            self._pending_src = self.real_AB_aug
            if self._sync_forward:
                torch.cuda.current_stream(self.device).synchronize()    # stats_time = wall time of the batch, as the reference's
            time_elapsed = time.time() - since
            if self.verbose_log:
                print('Augmentation completed in {:.0f}m {:.3f}s'.format(time_elapsed // 60, time_elapsed % 60))
        else:
            self._passthrough = True
            self.real_AB_aug = torch.cat((self.real_A, self.real_B), dim=1)
            self._pending, self._pending_src = (self.real_AB_aug, None), self.real_AB_aug
            time_elapsed = time.time() - since
            if self.verbose_log:
                print('No augmentation, time {:.0f}m {:.3f}s'.format(time_elapsed // 60, time_elapsed % 60))
        self.stats_time.append(time_elapsed)

    def iterate(self, loader, with_latents=False):
        """Look-ahead form of the caller loop (reference backbone_latentaug.py:91-124: ``for data in dataset:
        set_input; forward; get_output; <write>``): yields ``(data, output_dict)`` per batch with the forward of batch
        t+1 ENQUEUED before the output of batch t is awaited, so the device->host copy and the caller's work on batch
        t (pickling, disk writes) overlap the next batch's kernels.  Same values as the three-call sequence.
        ``with_latents``: yields ``(data, output_dict, get_latent_input() dict, get_latent_output() dict)`` of that batch --
        the two latent tensors travel on the copy stream too (calling ``get_latent_*`` inside the loop would read batch
        t+1's codes and wait for its kernels)."""
        prev = None
        self._sync_forward = False
        try:
            for data in loader:
                self.set_input(data)
                self.forward()
                cur = (data, self._pending, self.fname)
                if with_latents:
                    cur += (self._latents_d2h(),)
                if prev is not None:
                    yield self._yield(prev, with_latents)
                prev = cur
            if prev is not None:
                yield self._yield(prev, with_latents)
        finally:
            self._sync_forward = True

    def _latents_d2h(self):
        """Async copies of this batch's input / augmented codes (``None`` on the pass-through branch, where the reference's
        ``get_latent_*`` would read stale or missing attributes)."""
        if self._passthrough:
            return None
        return (self._start_d2h(self.w_AB, 'w_in'), self._start_d2h(reverse_broadcasting(self.w_AB_aug).contiguous(), 'w_aug'))

    def _yield(self, item, with_latents):
        out = self._finish(item[:3])
        if not with_latents:
            return item[0], out
        lat, fname = item[3], item[2]
        if lat is None:
            return item[0], out, None, None
        host = []
        for buf, ev in lat:
            if ev is not None:
                ev.synchronize()
            host.append(buf.numpy().squeeze().copy())          # owned, like the reference's .cpu().numpy()
        paths = fname if not self.rand_aug else ''
        return item[0], out, {'w': host[0], 'paths': paths}, {'w': host[1], 'paths': paths}

    def _finish(self, item):
        _, (buf, ev), fname = item
        if ev is not None:
            ev.synchronize()
        a = buf[:, 0, :, :].unsqueeze(dim=1)
        b = buf[:, min(1, buf.shape[1] - 1), :, :].unsqueeze(dim=1)
        if self.lower_bound_clip:
            a, b = torch.clamp(a, min=-1.0, max=None), torch.clamp(b, min=-1.0, max=None)
        return {'A': a, 'B': b, 'A_paths': fname, 'B_paths': fname}

    def sanity_check(self):
        """Smoke run of one batch (reference :281-301 also dumps PNGs with matplotlib; not reproduced)."""
        self.forward()
        data = self.get_output()
        res = self.opt.img_resolution
        assert data['A'].dtype == torch.float32 and data['A'].shape[1:] == (1, res, res)
        assert data['B'].dtype == torch.float32 and data['B'].shape[1:] == (1, res, res)

    def sample_from_randn(self):
        return torch.randn([self.batch_size, self.z_dim])

    def sample_from_inversion(self, fname):
        w = self.stats_dataset_w.lookup(fname).reshape(len(fname), 1, self.w_dim)
        assert w.shape == (self.batch_size, 1, self.w_dim)
        return w
