"""``AugOptions`` (reference ``options/aug_options.py:4-17``): the shared flags of ``BaseOptions`` plus ``--phase``."""
from .base_options import BaseOptions


class AugOptions(BaseOptions):
    def initialize(self, parser):
        parser = super().initialize(parser)
        parser.add_argument('--phase', default='train', type=str, help="'train', 'val' or 'test': which split of the inverted codes is read")
        self.isTrain = True
        return parser
