"""Generate tests/golden/formats/ -- two small zips in the reference's dataset layout and what the REFERENCE's own
``LatentCodeDataset`` / ``ImgDataset`` / ``DatasetStats`` / ``compute_stats`` filter make of them (build container only).

    python -m oracle.make_golden_formats
"""
import os
import pickle
import sys
import zipfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from oracle.ref_driver import REF_SRC, import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'formats')
NUM_WS, W_DIM, RES = 6, 16, 8
MODS = ['MR_nonrigid_CT', 'MR_MR_T2']


def write_zips():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.RandomState(3)
    with zipfile.ZipFile(os.path.join(OUT, 'codes.zip'), 'w') as zw, zipfile.ZipFile(os.path.join(OUT, 'images.zip'), 'w') as zi:
        for split in ('train', 'val'):
            for patient in ('Pelvis_2101', 'Pelvis_2102'):
                for sl in (5, 10, 12, 15, 20, 25, 120, 121):
                    name = f'{split}/{patient}/{patient}_{sl:05d}.pickle'        # data/write_tozip.py:41-46
                    w = np.repeat(rng.randn(1, W_DIM).astype('float32'), NUM_WS, axis=0)
                    zw.writestr(name, pickle.dumps(w, pickle.HIGHEST_PROTOCOL))
                    img = {m: rng.randint(0, 256, size=(RES, RES)).astype('uint8') for m in MODS}
                    zi.writestr(name, pickle.dumps(img, pickle.HIGHEST_PROTOCOL))


def main():
    write_zips()
    ref = import_reference(REF_SRC)
    uds = ref.uds
    ds_w = uds.LatentCodeDataset(os.path.join(OUT, 'codes.zip'), split='train', w_dim=W_DIM, num_ws=NUM_WS)
    ds_i = uds.ImgDataset(os.path.join(OUT, 'images.zip'), split='train', modalities=MODS, resolution=RES)
    out = {'fnames': list(ds_w._fnames), 'w0': torch.from_numpy(ds_w[0][0]), 'img3': torch.from_numpy(ds_i[3][0])}
    for step, manifold, ds in ((5, 'latent', ds_w), (10, 'img', ds_i)):
        stats = uds.DatasetStats(manifold=manifold, max_items=len(ds), step=step)
        loader = torch.utils.data.DataLoader(dataset=ds, batch_size=1, shuffle=False)        # util_latent_aug.py:506
        for x, fname in loader:
            if manifold == 'img':
                x = x / 127.5 - 1                                                            # :544
            if stats.append_torch(x, fname) < 0:
                break
        out[f'{manifold}_step{step}'] = stats.get_all_torch()
        out[f'{manifold}_schedule{step}'] = list(stats.schedule)
        stats.save(os.path.join(OUT, f'ref_cache_{manifold}.pkl'))
    torch.save(out, os.path.join(OUT, 'expected.pt'))
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == '__main__':
    main()
