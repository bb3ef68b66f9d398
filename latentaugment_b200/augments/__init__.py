"""Augment plugin registry -- same surface as the reference ``augments/__init__.py:28-72``:
``find_augment_using_name``, ``get_option_setter``, ``create_augment``."""
import importlib

from .base_aug import BaseAugment


def find_augment_using_name(augment_name):
    """Import ``augments/[augment_name]_aug.py`` and return the BaseAugment subclass whose
    lower-cased name is ``[augmentname]augment`` (reference augments/__init__.py:28-48)."""
    augment_filename = __name__ + '.' + augment_name + '_aug'
    augmentlib = importlib.import_module(augment_filename)
    augment = None
    target_augment_name = augment_name.replace('_', '') + 'augment'
    for name, cls in augmentlib.__dict__.items():
        if name.lower() == target_augment_name.lower() and isinstance(cls, type) and issubclass(cls, BaseAugment):
            augment = cls
    if augment is None:
        raise ImportError('In %s.py, there should be a subclass of BaseAugment with class name that matches %s in lowercase.'
                          % (augment_filename, target_augment_name))
    return augment


def get_option_setter(augment_name):
    """Static method ``modify_commandline_options`` of the augment class (reference :51-54)."""
    return find_augment_using_name(augment_name).modify_commandline_options


def create_augment(opt):
    """``augment = create_augment(opt)`` (reference :57-72)."""
    augment = find_augment_using_name(opt.aug)
    instance = augment(opt)
    print('Augment [%s] was created' % type(instance).__name__)
    return instance
