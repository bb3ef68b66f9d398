"""Abstract base class of augment plugins (reference ``augments/base_aug.py:7-64``): what ``create_augment`` hands back and
what the training loop calls -- ``set_input(data)`` then ``forward()``; the three hooks below are optional."""
import os
from abc import ABC, abstractmethod

import torch


class BaseAugment(ABC):
    def __init__(self, opt):
        self.opt, self.gpu_ids = opt, opt.gpu_ids
        self.device = torch.device(f'cuda:{self.gpu_ids[0]}' if self.gpu_ids else 'cpu')      # where the CALLER's tensors live
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)

    @staticmethod
    def modify_commandline_options(parser, is_train):
        """A plugin adds its own flags to ``parser`` here (options/base_options.py gathers them); the base adds none."""
        return parser

    @abstractmethod
    def set_input(self, data):
        """Takes one batch dict from the data loader."""

    @abstractmethod
    def forward(self):
        """Augments the batch given to ``set_input``."""

    # optional hooks of the reference interface: no-ops unless a plugin overrides them
    def get_train_transform(self):
        return None

    def get_valid_transform(self):
        return None

    def sanity_check(self):
        return None
