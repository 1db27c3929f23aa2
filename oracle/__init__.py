"""CPU oracle for the RoViT-KAN forward/backward path.

TEST INFRASTRUCTURE ONLY.  This package restates, in plain fp32 PyTorch/numpy on
the CPU, the arithmetic the reference performs on the hot path
(`models/{kan,heads,backbone,rovit_kan}.py`, `training/losses.py` of the
reference, plus the `timm` DeiT-Tiny trunk that the reference pulls in as a
third-party dependency, `timm>=0.6.0`, un-pinned and un-vendored).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import it, and only as the checker.  The
product path (`rovitkan_b200`) never imports this package and fails loudly when
its CUDA library is missing.

Pinning status
--------------
* KAN, heads, losses: PINNED.  `tests/golden/make_golden.py` imports the
  reference's own `models/kan.py`, `models/heads.py`, `training/losses.py` in
  the build container, runs them on seeded inputs and stores the outputs (and
  autograd gradients) in `tests/golden/*.npz`; `tests/test_oracle_golden.py`
  checks every oracle function against those vectors plus the RNG-free
  known-answer tables of SURVEY.md section 4.
* DeiT-Tiny trunk: the reference holds no vectors for it and `timm` is absent
  from the image, so this part is "parity unpinned" against timm itself.  It is
  pinned instead against two independent implementations of the same published
  architecture that ARE in the image (torchvision `VisionTransformer`,
  HuggingFace `ViTModel`) through weight remapping; golden vectors from those
  runs are committed too.
"""

from . import kan, vit, heads, losses, model  # noqa: F401
