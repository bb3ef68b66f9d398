"""One-shot converter for the reference's network pickles (``network-snapshot-*.pkl``, loaded by the reference at
``augments/utils/util_latent_aug.py:466-484``): dumps ``G_ema`` and ``D`` as plain ``state_dict`` files that
``--generator_state`` / ``--discriminator_state`` take.

    python tools/export_reference_pickle.py network-snapshot-005320.pkl --out-dir states/ [--reference /path/to/LatentAugment]

The pickles embed the source of the network classes (``torch_utils/persistence.py:118-126,179-227``) and need the
reference's ``torch_utils`` / ``dnnlib`` packages importable to be unpickled: pass ``--reference`` (default: the copy
under ``baseline/_ref`` if present).  Unpickling executes the embedded source -- only convert files you trust.
The parameter names written are exactly the ones the engine reads (``models/stylegan3/legacy.py:171-203,267-287``).
"""
import argparse
import os
import pickle
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('network_pkl')
    ap.add_argument('--out-dir', default='.')
    ap.add_argument('--reference', default=os.path.join(ROOT, 'baseline', '_ref'))
    a = ap.parse_args()
    sg3 = os.path.join(a.reference, 'models', 'stylegan3')
    if not os.path.isdir(os.path.join(sg3, 'torch_utils')):
        sys.exit(f'{sg3}/torch_utils not found: pass --reference <checkout of the reference repository>')
    sys.path[:0] = [a.reference, sg3]
    import torch
    with open(a.network_pkl, 'rb') as f:
        nets = pickle.load(f)
    os.makedirs(a.out_dir, exist_ok=True)
    for key, fname in (('G_ema', 'generator_state.pt'), ('D', 'discriminator_state.pt')):
        if key not in nets:
            continue
        sd = {k: v.detach().cpu().float() for k, v in nets[key].state_dict().items()}
        torch.save(sd, os.path.join(a.out_dir, fname))
        print(f'{key}: {len(sd)} tensors -> {os.path.join(a.out_dir, fname)}')


if __name__ == '__main__':
    main()
