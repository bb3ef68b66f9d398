"""CPU oracle for the LatentAugment latent-optimisation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``latentaugment_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker or as
the CPU baseline being timed -- never as the product path.

What it is: a plain-torch (CPU, fp32/fp64) restatement of the reference's algorithm
for the path ``LatentAugment.forward`` -> ``LatentAug.forward`` -> ``G.synthesis``
fwd/bwd-to-w -> criteria losses -> Adam.  Every function cites the reference
file:line it follows (paths relative to ``/root/reference``).

Pinning status
--------------
* The reference ships NO tests / golden vectors, so the pins are
  manufactured by running the reference's own code in the build container
  (``oracle/make_golden.py``) and committing the outputs under ``tests/golden/``:
  op level (``_bias_act_ref``, ``_upfirdn2d_ref``, ``upsample2d``,
  ``conv2d_resample``, ``fma``), loss level (``LatentAug.l2_loss_vectorized``,
  ``calc_loss_latent``, ``calc_loss_pix``) and loop level (the reference's
  ``LatentAug.forward`` itself, driven through a generator assembled from the
  reference's ``torch_utils.ops``).
* The StyleGAN2 network classes are NOT in the reference tree:
  they are unpickled from third-party source (NVlabs/stylegan3
  ``training/networks_stylegan2.py``, no version pin exists in the reference).
  ``oracle/sg2.py`` restates that published architecture against the in-tree
  parameter contract ``models/stylegan3/legacy.py:122-203``.  The generator
  *architecture* is therefore pinned only structurally ("parity unpinned" for the
  network classes themselves); every op it is composed of, and the whole
  optimisation loop around it, is pinned against reference outputs.
"""
