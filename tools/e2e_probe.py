"""Times set_input / forward / get_output of the plugin separately per iteration (GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from latentaugment_b200.augments import create_augment
from latentaugment_b200.options.aug_options import AugOptions

B, res = 32, 256
argv = ['--aug', 'latent', '--synthetic', '--batch_size', str(B), '--gpu_ids', '0', '--gpu_ids_aug', '0', '--img_resolution', str(res),
        '--synthetic_channels', '3', '--synthetic_bank', '4096', '--synthetic_img_bank', '64', '--synthetic_codes', '256',
        '--precision', 'bf16', '--opt_num_epochs', '10', '--no_log']
opt = AugOptions().parse(args={'p_thres': 0.0, 'w_lpips': 0.0, 'w_disc': 0.0, 'init_w': 'inv', 'n_imgs': 0}, argv=argv)
aug = create_augment(opt)
names = list(aug.stats_dataset_w.index.keys())
img = torch.zeros([B, 1, res, res])
for i in range(12):
    fn = [names[(i * B + j) % len(names)] for j in range(B)]
    t0 = time.perf_counter()
    aug.set_input({'A': img, 'B': img, 'A_paths': fn, 'B_paths': fn})
    t1 = time.perf_counter()
    aug.forward()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    out = aug.get_output()
    t4 = time.perf_counter()
    print(f'iter {i}: set_input {1e3*(t1-t0):.2f} ms, forward(launch) {1e3*(t2-t1):.2f} ms, gpu wait {1e3*(t3-t2):.2f} ms, get_output {1e3*(t4-t3):.2f} ms', file=sys.stderr)
