"""Times every tap-GEMM launch of one Adam step alone (GPU box).  With --reps 1 each launch runs three times in
the order fwd(l) x3, dgrad(l) x3 for l = 0..L-1 -- the fixed sequence the ncu captures under profiles/ index into
(`ncu -k regex:tapgemm_kernel --launch-skip 6*l+...`).  LA_DBG_CLK=1 adds the in-kernel timeline of CTA 0."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from latentaugment_b200.augments import create_augment
from latentaugment_b200.options.aug_options import AugOptions

ap = argparse.ArgumentParser()
ap.add_argument('--reps', type=int, default=10)
ap.add_argument('--precision', default='bf16')
ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--res', type=int, default=256)
a = ap.parse_args()
argv = ['--aug', 'latent', '--synthetic', '--batch_size', str(a.batch), '--gpu_ids', '0', '--gpu_ids_aug', '0', '--img_resolution', str(a.res),
        '--synthetic_channels', '3', '--synthetic_bank', '4096', '--synthetic_img_bank', '64', '--synthetic_codes', '256',
        '--precision', a.precision, '--opt_num_epochs', '10', '--no_log']
opt = AugOptions().parse(args={'p_thres': 0.0, 'w_lpips': 0.0, 'w_disc': 0.0, 'init_w': 'inv', 'n_imgs': 0}, argv=argv)
aug = create_augment(opt)
eng = aug.latent_aug.module.engines[0]
names = list(aug.stats_dataset_w.index.keys())
img = torch.zeros([a.batch, 1, a.res, a.res])
fn = [names[j % len(names)] for j in range(a.batch)]
aug.set_input({'A': img, 'B': img, 'A_paths': fn, 'B_paths': fn})
aug.forward()                      # fills every activation buffer the launches read
torch.cuda.synchronize()
torch.cuda.nvtx.range_push('gemms')        # ncu --nvtx --nvtx-include gemms/ sees only the timed launches
t = eng.debug_time_gemms(reps=a.reps)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
for k in ('forward', 'dgrad', 'fir_forward', 'fir_backward'):
    if k in t:
        print(k, ' '.join(f'{v:.3f}' for v in t[k]), f'| sum {sum(t[k]):.3f} ms')
print({k: v for k, v in t.items() if not isinstance(v, (list, tuple))})
