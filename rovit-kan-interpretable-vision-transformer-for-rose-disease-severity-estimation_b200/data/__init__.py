"""Stand-in for the reference's `data` package, which the reference git-ignores (.gitignore:60) although its
trainer and every script import it (training/trainer.py:9, scripts/train.py:12-13, scripts/evaluate.py:10-11).

Surface inferred from those call sites: `data.dataset.{RoseLeafDataset, create_dataloaders}`,
`data.transforms.{augmented_transforms, original_transforms, inference_transforms, cutmix_or_mixup}`.
Image folders are read when they exist; otherwise the dataset is synthetic (seeded, class-dependent patterns),
which is what the benchmarks and the drop-in tests use.  Outside the hot path: plain torch, no kernels here.
"""

from . import dataset, transforms  # noqa: F401
