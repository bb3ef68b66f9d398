"""The reference's on-disk formats (zip of pickles, slice-schedule filter, cache pickle) read by the product's
``augments/utils/util_dataset.py`` against what the REFERENCE's own classes made of the same zips
(tests/golden/formats/*, oracle/make_golden_formats.py).  CPU only."""
import os
import pickle
import types

import pytest
import torch

from conftest import GOLDEN

F = os.path.join(GOLDEN, 'formats')
MODS = ['MR_nonrigid_CT', 'MR_MR_T2']


@pytest.fixture(scope='module')
def expected():
    return torch.load(os.path.join(F, 'expected.pt'), weights_only=True)


def test_zip_readers_match_reference(expected):
    from latentaugment_b200.augments.utils import util_dataset as ud
    ds_w = ud.LatentCodeDataset(os.path.join(F, 'codes.zip'), split='train', w_dim=16, num_ws=6)
    ds_i = ud.ImgDataset(os.path.join(F, 'images.zip'), split='train', modalities=MODS, resolution=8)
    assert ds_w.fnames == expected['fnames'] and len(ds_w) == 16 and len(ds_i) == 16
    w, name = ds_w[0]
    assert name == expected['fnames'][0] and torch.equal(torch.from_numpy(w), expected['w0'])
    assert torch.equal(torch.from_numpy(ds_i[3][0]), expected['img3'])
    with pytest.raises(IOError):
        ud.LatentCodeDataset(os.path.join(F, 'codes.zip'), split='train', w_dim=32, num_ws=6)
    with pytest.raises(IOError):
        ud.ImgDataset(os.path.join(F, 'images.zip'), split='test', modalities=MODS, resolution=8)       # no such split


def test_slice_schedule_filter_and_banks_match_reference(expected, tmp_path):
    from latentaugment_b200.augments.utils import util_dataset as ud
    ds_w = ud.LatentCodeDataset(os.path.join(F, 'codes.zip'), split='train', w_dim=16, num_ws=6)
    ds_i = ud.ImgDataset(os.path.join(F, 'images.zip'), split='train', modalities=MODS, resolution=8)
    sw = ud.compute_stats(ds_w, 'latent', str(tmp_path), step=5)
    si = ud.compute_stats(ds_i, 'img', str(tmp_path), step=10)
    assert sw.schedule == expected['latent_schedule5'] and si.schedule == expected['img_schedule10']
    assert torch.equal(sw.get_all_torch(), expected['latent_step5'])          # slices 10, 15, 20, 25, 120 of both patients
    assert torch.equal(si.get_all_torch(), expected['img_step10'])            # slices 10, 20, 120; images in [-1, 1]
    assert sw.get_all_torch().shape == (10, 6, 16) and si.get_all_torch().shape == (6, 2, 8, 8)
    # cache: written with the reference's tag, read back identically; the reference's own cache file loads too
    assert os.path.isfile(tmp_path / 'latent-step_5-maxitems_16.pkl')
    again = ud.compute_stats(ds_w, 'latent', str(tmp_path), step=5)
    assert torch.equal(again.get_all_torch(), expected['latent_step5'])
    ref_cache = ud.DatasetStats.load(os.path.join(F, 'ref_cache_img.pkl'))
    assert ref_cache.manifold == 'img' and torch.equal(ref_cache.get_all_torch(), expected['img_step10'])


def test_inverted_code_table_from_zip(expected):
    from latentaugment_b200.augments.utils import util_dataset as ud
    ds_w = ud.LatentCodeDataset(os.path.join(F, 'codes.zip'), split='train', w_dim=16, num_ws=6)
    table = ds_w.to_table()
    names = [expected['fnames'][5], expected['fnames'][0]]
    got = table.lookup(names)
    assert got.shape == (2, 16)
    assert torch.equal(got[1], expected['w0'][0]) and torch.equal(got[0], torch.from_numpy(ds_w[5][0][0]))


def test_reference_directory_layout(expected, tmp_path):
    """{interim_dir}/{dataset_aug}/{dataset_w_name}.zip + {dataset_name_aug}.zip, caches under cache_dir (util_latent_aug.py:133-158)."""
    import shutil

    from latentaugment_b200.augments.utils import util_dataset as ud
    root = tmp_path / 'interim' / 'Pelvis'
    os.makedirs(root)
    shutil.copy(os.path.join(F, 'codes.zip'), root / 'codes-w.zip')
    shutil.copy(os.path.join(F, 'images.zip'), root / 'imgs.zip')
    opt = types.SimpleNamespace(interim_dir=str(tmp_path / 'interim'), dataset_aug='Pelvis', dataset_w_name='codes-w', dataset_name_aug='imgs',
                                modalities_aug=','.join(MODS), img_resolution=8, step_w=5, step_img=10)
    ds_w, W, X = ud.banks_from_reference_layout(opt, 'train', 16, 6)
    assert torch.equal(W, expected['latent_step5']) and torch.equal(X, expected['img_step10'])
    assert os.path.isfile(root / 'cache_dir' / 'img-step_10-maxitems_16.pkl')


def test_unpickler_refuses_code_execution(tmp_path):
    from latentaugment_b200.augments.utils import util_dataset as ud

    class Evil:
        def __reduce__(self):
            return (os.system, ('true',))
    p = tmp_path / 'evil.pkl'
    with open(p, 'wb') as f:
        pickle.dump(Evil(), f)
    with open(p, 'rb') as f, pytest.raises(pickle.UnpicklingError):
        ud.safe_pickle_load(f)
