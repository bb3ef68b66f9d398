"""Bank / inverted-code I/O in the reference's own on-disk formats (reference ``augments/utils/util_dataset.py:35-279``,
``augments/utils/util_latent_aug.py:503-563`` ``compute_stats``, ``augments/latent_aug.py:310-324``; writers
``data/write_tozip.py:30-68``):

* latent zip  ``{dataset_w_name}.zip``: entries ``{split}/{patient}/{patient}_{ddddd}.pickle`` = pickled ``ndarray [num_ws, w_dim]``;
* image zip   ``{dataset_name}.zip``:   entries ``{split}/{patient}/{patient}_{ddddd}.pickle`` = pickled
  ``dict[modality -> HxW array in 0..255]``;
* bank = the entries whose 5-digit slice id is in the schedule ``{10, 10+step, ..., <= 120}`` (``DatasetStats``, :45,78-84),
  at most ``max_items``, images mapped to ``x / 127.5 - 1``; cached as a pickle of ``DatasetStats.__dict__`` under
  ``cache_dir/{tag}.pkl`` with the reference's tag scheme so caches are interchangeable.

What changes against the reference is only WHERE the data lives afterwards: the inverted codes go into ONE pinned
``[N, w_dim]`` table (``InvertedCodeTable``; the reference re-opens the zip and unpickles per sample on every batch) and
the banks go to the GPU once as bank moments.  Unpickling is restricted to numpy arrays / plain containers.
"""
import io
import os
import pickle
import zipfile

import numpy as np
import torch

MAX_ITEMS = 100000          # compute_stats default (util_latent_aug.py:503)


class _NumpyUnpickler(pickle.Unpickler):
    """The zips hold pickled numpy arrays (and dicts of them): nothing else is allowed to be constructed."""
    _ALLOWED = {('numpy', 'ndarray'), ('numpy', 'dtype'), ('numpy.core.multiarray', '_reconstruct'), ('numpy._core.multiarray', '_reconstruct'),
                ('numpy.core.multiarray', 'scalar'), ('numpy._core.multiarray', 'scalar'), ('numpy.core.numeric', '_frombuffer'),
                ('numpy._core.numeric', '_frombuffer'), ('collections', 'OrderedDict'), ('builtins', 'dict'), ('builtins', 'list'),
                ('builtins', 'tuple'), ('builtins', 'set'), ('builtins', 'str'), ('builtins', 'int'), ('builtins', 'float'),
                ('builtins', 'bool'), ('builtins', 'NoneType'), ('builtins', 'slice')}

    def find_class(self, module, name):
        if (module, name) in self._ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f'refusing to unpickle {module}.{name}: the dataset formats hold numpy arrays and plain containers only')


def safe_pickle_load(f):
    return _NumpyUnpickler(f).load()


def slice_schedule(step):
    """util_dataset.py:45 -- the 5-digit slice ids kept per patient."""
    return sorted(f'{i:05d}' for i in np.arange(start=10, stop=120 + 1, step=step))


def slice_id(fname):
    """util_dataset.py:78-79 -- last five characters of the file name without extension."""
    return os.path.splitext(os.path.basename(fname))[0][-5:]


class _ZipPickleDataset(torch.utils.data.Dataset):
    def __init__(self, path, split):
        self._path, self._split, self._zipfile = path, split, None
        if os.path.splitext(path)[1].lower() != '.zip':
            raise IOError('Path must point to a zip')
        names = set(self._get_zipfile().namelist())
        self._fnames = sorted(f for f in names if os.path.splitext(f)[1].lower() == '.pickle' and split in f)
        if not self._fnames:
            raise IOError('No files found in the specified path')

    def _get_zipfile(self):
        if self._zipfile is None:
            self._zipfile = zipfile.ZipFile(self._path)
        return self._zipfile

    def open_file(self, fname):
        return self._get_zipfile().open(fname, 'r')

    def __len__(self):
        return len(self._fnames)

    @property
    def fnames(self):
        return list(self._fnames)

    def __getstate__(self):
        return dict(self.__dict__, _zipfile=None)


class LatentCodeDataset(_ZipPickleDataset):
    """util_dataset.py:150-210: ``ds[i] -> (w [num_ws, w_dim] float32, entry name)``."""

    def __init__(self, path, split, w_dim=512, num_ws=14):
        super().__init__(path, split)
        shape = self._load_w(0)[0].shape
        if w_dim is not None and shape[1] != w_dim:
            raise IOError('W does not match the specified latent dimension.')
        if num_ws is not None and shape[0] != num_ws:
            raise IOError('W does not match the specified broadcasting.')
        self._raw_shape = [len(self._fnames)] + list(shape)

    def _load_w(self, raw_idx):
        fname = self._fnames[raw_idx]
        with self.open_file(fname) as f:
            w = safe_pickle_load(io.BytesIO(f.read()))
        return np.asarray(w).astype('float32'), fname

    def __getitem__(self, idx):
        return self._load_w(idx)

    def to_table(self):
        """All inverted codes as one ``InvertedCodeTable`` keyed by entry name (row 0 of each ``[num_ws, w_dim]`` code:
        ``reverse_broadcasting``, latent_aug.py:321) -- read once instead of per sample per batch."""
        from .util_latent_aug import InvertedCodeTable
        codes = torch.from_numpy(np.stack([self._load_w(i)[0][0] for i in range(len(self))]))
        return InvertedCodeTable(self.fnames, codes)


class ImgDataset(_ZipPickleDataset):
    """util_dataset.py:212-279: ``ds[i] -> (image [C, H, W] float32 in 0..255, entry name)``; channel order = ``modalities``."""

    def __init__(self, path, split, modalities, resolution=256):
        super().__init__(path, split)
        self._modalities = list(modalities)
        assert len(self._modalities) > 0
        shape = self._load_raw_image(0)[0].shape
        if resolution is not None and (shape[1] != resolution or shape[2] != resolution):
            raise IOError('Image files do not match the specified resolution')
        self._raw_shape = [len(self._fnames)] + list(shape)

    def _load_raw_image(self, raw_idx):
        fname = self._fnames[raw_idx]
        with self.open_file(fname) as f:
            p = safe_pickle_load(io.BytesIO(f.read()))
        s = np.asarray(p[self._modalities[0]])
        out = np.zeros((len(self._modalities), s.shape[0], s.shape[1]), dtype='float32')
        for i, m in enumerate(self._modalities):
            out[i] = np.asarray(p[m]).astype('float32')
        return out, fname

    def __getitem__(self, idx):
        return self._load_raw_image(idx)


class DatasetStats:
    """util_dataset.py:35-147 (same attributes, so ``save`` / ``load`` caches are interchangeable with the reference's)."""
    _NDIM = {'latent': 3, 'features': 4, 'features_jit': 2, 'img': 4}

    def __init__(self, manifold, capture_all=False, max_items=None, step=1):
        if manifold not in self._NDIM:
            raise NotImplementedError('Unrecognised manifold! Add it!')
        self.manifold, self.capture_all, self.max_items, self.step = manifold, capture_all, max_items, step
        self.num_items = 0
        self.all_x = []
        self.schedule = slice_schedule(step)
        self.ndim = self._NDIM[manifold]

    def is_full(self):
        return self.max_items is not None and self.num_items >= self.max_items

    def append(self, x, fname):
        """``x`` [1, ...]; ``fname`` = the entry name (or a 1-list of it, as the reference's DataLoader delivers).  Returns
        the number of rows added, 0 if the slice is off-schedule, -1 when full."""
        x = np.asarray(x, dtype=np.float32)
        assert x.ndim == self.ndim
        if self.max_items is not None and self.num_items + x.shape[0] > self.max_items:
            if self.num_items >= self.max_items:
                return -1
            x = x[:self.max_items - self.num_items]
        if not self.capture_all:
            name = fname[0] if isinstance(fname, (list, tuple)) else fname
            if slice_id(name) not in self.schedule:
                return 0
        self.all_x.append(x)
        self.num_items += x.shape[0]
        return x.shape[0]

    def append_torch(self, x, idd=None):
        assert isinstance(x, torch.Tensor) and x.ndim == self.ndim and x.shape[0] == 1
        return self.append(x.cpu().numpy(), idd)

    def get_all(self):
        return np.concatenate(self.all_x, axis=0)

    def get_all_torch(self):
        return torch.from_numpy(self.get_all().astype(np.float32))

    def save(self, pkl_file):
        with open(pkl_file, 'wb') as f:
            pickle.dump(self.__dict__, f)

    @staticmethod
    def load(pkl_file):
        with open(pkl_file, 'rb') as f:
            s = safe_pickle_load(f)
        obj = DatasetStats(manifold=s['manifold'], capture_all=s['capture_all'], max_items=s['max_items'], step=s['step'])
        obj.__dict__.update(s)
        return obj


def cache_tag(manifold, step, num_items, tag=''):
    """util_latent_aug.py:517-523"""
    base = f'{manifold}-step_{step}-maxitems_{num_items}'
    return f'{tag}-{base}' if tag else base


def compute_stats(dataset, manifold, cache_dir=None, tag='', step=10, max_items=MAX_ITEMS):
    """util_latent_aug.py:503-563 for the 'latent' and 'img' manifolds: walks the dataset in entry order, keeps the
    on-schedule slices (images mapped to [-1, 1], :544), loads / writes the cache pickle.  (The feature manifolds are not
    cached as tensors here: the engine builds its feature-bank moments from the image bank, criteria/lpips.py.)"""
    if manifold not in ('latent', 'img'):
        raise NotImplementedError(manifold)
    num_items = len(dataset) if max_items is None else min(len(dataset), max_items)
    cache_file = os.path.join(cache_dir, cache_tag(manifold, step, num_items, tag) + '.pkl') if cache_dir else None
    if cache_file and os.path.isfile(cache_file):
        print(f'{manifold} dataset already created in {cache_file}.')
        return DatasetStats.load(cache_file)
    print(f'{manifold} dataset initialization.')
    stats = DatasetStats(manifold=manifold, max_items=num_items, step=step)
    for i in range(len(dataset)):
        name = dataset._fnames[i]
        if slice_id(name) not in stats.schedule:          # decided by the NAME: off-schedule entries are never unpickled
            continue
        x, fname = dataset[i]
        x = np.asarray(x)[None]
        if manifold == 'img':
            x = x / 127.5 - 1
        if stats.append(x, fname) < 0:
            break
    if cache_file:
        os.makedirs(cache_dir, exist_ok=True)
        stats.save(cache_file)
    return stats


def banks_from_reference_layout(opt, phase, w_dim, num_ws, need_latent=True, need_img=True):
    """The reference's directory convention (util_latent_aug.py:133-158):
    ``{interim_dir}/{dataset_aug}/{dataset_w_name}.zip`` (inverted codes, split = phase) and
    ``{interim_dir}/{dataset_aug}/{dataset_name_aug}.zip`` (images), caches under ``.../cache_dir``.
    Returns ``(inverted-code dataset, latent bank [M, num_ws, w_dim] or None, image bank [M, C, res, res] or None)``."""
    root = os.path.join(opt.interim_dir, opt.dataset_aug)
    cache_dir = os.path.join(root, 'cache_dir')
    ds_w = LatentCodeDataset(os.path.join(root, opt.dataset_w_name + '.zip'), split=phase, w_dim=w_dim, num_ws=num_ws)
    W = X = None
    if need_latent:
        W = compute_stats(ds_w, 'latent', cache_dir, step=opt.step_w).get_all_torch()
    if need_img:
        mods = [m for m in str(opt.modalities_aug).split(',') if m]
        ds_i = ImgDataset(os.path.join(root, opt.dataset_name_aug + '.zip'), split=phase, modalities=mods, resolution=opt.img_resolution)
        X = compute_stats(ds_i, 'img', cache_dir, step=opt.step_img).get_all_torch()
    return ds_w, W, X
