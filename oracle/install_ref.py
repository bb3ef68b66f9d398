"""Recipe that places an UNMODIFIED copy of the reference's hot-path files under ``baseline/_ref`` (git-ignored, so
the history stays free of reference sources; NOT gpurun-ignored, so the copy travels to the GPU box where
``/root/reference`` does not exist).  MEASUREMENT INFRASTRUCTURE: only bench.py's ``cpu_baseline`` / ``gpu_reference`` /
``--impl reference`` legs read it (through oracle/ref_driver.py).

    python -m oracle.install_ref            # no-op when /root/reference is absent (GPU box) or the copy is current

The reference has no setup.py / pyproject.toml, so ``pip install --target baseline/_ref /root/reference`` is not
applicable; a plain copy of the directories the path imports is the install.
"""
import os
import shutil

from .ref_driver import REF_COPY, REF_SRC

# what augments/utils/util_latent_aug.py and torch_utils/ops import, transitively
TREES = ['augments', 'utils', 'options', 'models/stylegan3/torch_utils', 'models/stylegan3/dnnlib']
FILES = ['__init__.py', 'models/stylegan3/legacy.py', 'README.md']


def install(src=REF_SRC, dst=REF_COPY, force=False):
    if not os.path.isdir(src):
        return None
    stamp = os.path.join(dst, '.installed')
    if os.path.exists(stamp) and not force:
        return dst
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    for t in TREES:
        shutil.copytree(os.path.join(src, t), os.path.join(dst, t), ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    for f in FILES:
        if os.path.exists(os.path.join(src, f)):
            os.makedirs(os.path.dirname(os.path.join(dst, f)) or dst, exist_ok=True)
            shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
    open(stamp, 'w').write('copied from %s\n' % src)
    return dst


if __name__ == '__main__':
    print(install(force=True))
