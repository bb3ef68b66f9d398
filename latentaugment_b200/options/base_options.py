"""Two-pass argparse option gathering (reference ``options/base_options.py:22-175``): base options,
then the options the selected augment plugin contributes; ``parse(args=dict)`` overrides."""
import argparse
import os
import sys

import torch

from .. import augments


class _Tee:
    """stdout tee to ``checkpoints_dir/name/log.txt`` (reference utils/util_logger.py:6-58)."""

    def __init__(self, file_name, file_mode='a'):
        self.file = open(file_name, file_mode)
        self.stdout = sys.stdout
        sys.stdout = self

    def write(self, text):
        if text:
            self.file.write(text)
            self.file.flush()
            self.stdout.write(text)

    def flush(self):
        self.file.flush()
        self.stdout.flush()

    def close(self):
        if sys.stdout is self:
            sys.stdout = self.stdout
        self.file.close()


class BaseOptions:
    def __init__(self):
        self.initialized = False
        self.isTrain = True

    def initialize(self, parser):
        parser.add_argument('--dataroot', default='', help='path to images (unused by the augmentation path itself)')
        parser.add_argument('--name', type=str, default='experiment_name')
        parser.add_argument('--gpu_ids', type=str, default='0')
        parser.add_argument('--checkpoints_dir', type=str, default='./checkpoints')
        parser.add_argument('--dataset_mode', type=str, default='pelvis2.1')
        parser.add_argument('--load_size', type=int, default=256)
        parser.add_argument('--aug', type=str, default=None, help='Augmentation mode [latent]')
        parser.add_argument('--batch_size', type=int, default=1)
        parser.add_argument('--serial_batches', action='store_true')
        parser.add_argument('--max_dataset_size', type=float, default=float('inf'))
        parser.add_argument('--verbose', action='store_true')
        parser.add_argument('--suffix', default='', type=str)
        parser.add_argument('--no_log', action='store_true', help='do not tee stdout / write the option dump')
        self.initialized = True
        return parser

    def gather_options(self, argv=None):
        parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
        parser = self.initialize(parser)
        opt, _ = parser.parse_known_args(argv)
        # dataset options: the reference asks data.get_option_setter(dataset_mode) here (:58-62); the
        # data loaders are caller-side and contribute no option the path reads.
        if opt.aug is not None:
            parser = augments.get_option_setter(opt.aug)(parser, self.isTrain)
            opt, _ = parser.parse_known_args(argv)
        self.parser = parser
        return parser.parse_args(argv)

    def print_options(self, opt):
        message = '----------------- Options ---------------\n'
        for k, v in sorted(vars(opt).items()):
            default = self.parser.get_default(k)
            comment = '\t[default: %s]' % str(default) if v != default else ''
            message += '{:>25}: {:<30}{}\n'.format(str(k), str(v), comment)
        message += '----------------- End -------------------'
        print(message)
        expr_dir = os.path.join(opt.checkpoints_dir, opt.name)
        os.makedirs(expr_dir, exist_ok=True)
        with open(os.path.join(expr_dir, '{}_opt.txt'.format(opt.phase)), 'wt') as f:
            f.write(message + '\n')

    def parse(self, args=None, argv=None):
        """``args``: dict of overrides exactly as the reference applies them (:104-140); ``argv``:
        optional argument list instead of ``sys.argv`` (addition, for programmatic use)."""
        opt = self.gather_options(argv)
        opt.n_imgs = getattr(opt, 'n_imgs', 0)
        if args is not None:
            keys = list(args.keys())
            if 'n_imgs' in keys:
                opt.n_imgs = args['n_imgs']
            if opt.aug == 'latent' and opt.rand_aug:
                for k in ('p_thres', 'truncation_psi'):
                    if k in keys:
                        setattr(opt, k, args[k])
            else:
                for k in ('p_thres', 'opt_num_epochs', 'opt_lr', 'w_lpips', 'w_pix', 'w_latent', 'w_disc', 'init_w'):
                    if k in keys:
                        setattr(opt, k, args[k])
        opt.isTrain = self.isTrain
        if opt.aug is not None:
            if opt.aug == 'latent' and opt.rand_aug:
                suffix = f'n_imgs_{opt.n_imgs}-truncation_psi_{opt.truncation_psi}'
            else:
                suffix = (f'n_imgs_{opt.n_imgs}-opt_lr_{opt.opt_lr}-opt_num_epochs_{opt.opt_num_epochs}-w_latent_{opt.w_latent}'
                          f'-w_pix_{opt.w_pix}-w_lpips_{opt.w_lpips}-w_disc_{opt.w_disc}')
            opt.name = opt.name + '-' + suffix
        if not opt.no_log:
            os.makedirs(os.path.join(opt.checkpoints_dir, opt.name), exist_ok=True)
            self.logger = _Tee(os.path.join(opt.checkpoints_dir, opt.name, 'log.txt'))
            self.print_options(opt)
        str_ids = opt.gpu_ids.split(',')
        opt.gpu_ids = [int(s) for s in str_ids if int(s) >= 0]
        if len(opt.gpu_ids) > 0 and torch.cuda.is_available():
            torch.cuda.set_device(opt.gpu_ids[0])
        self.opt = opt
        return self.opt
