"""``AugOptions`` (reference ``options/aug_options.py:4-17``)."""
from .base_options import BaseOptions


class AugOptions(BaseOptions):
    def initialize(self, parser):
        parser = BaseOptions.initialize(self, parser)
        parser.add_argument('--phase', type=str, default='train', help='train, val, test, etc')
        self.isTrain = True
        return parser
