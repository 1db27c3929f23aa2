"""`data.dataset` surface used by the reference (scripts/train.py:12,73-84,110-111; scripts/evaluate.py:10,40-54;
scripts/run_ablation.py:26-42,147-152): `RoseLeafDataset` with `.samples` / `.classes` / `.get_class_weights()` and
`create_dataloaders(...) -> (train, val, test)` whose train/val datasets are `Subset`s of one RoseLeafDataset.

Without image folders the dataset is synthetic, and its samples are REAL files: `scripts/run_ablation.py:34-42` and
`scripts/run_baselines.py` re-open `samples[i]['path']` with PIL themselves (to swap the transform of a Subset), so every
synthetic image is written once as a PNG under `$ROVITKAN_SYNTH_DIR` (default: `<tmp>/rovitkan_synthetic_<uid>`) and read back
from there by both routes."""

from __future__ import annotations

import os
import tempfile
import zlib
from typing import Dict, List, Optional

import torch
from torch.utils.data import DataLoader, Dataset, random_split

from .transforms import IMAGE_SIZE

_EXT = ('.jpg', '.jpeg', '.png', '.bmp', '.gif', '.webp')
DEFAULT_CLASSES = ['Healthy Leaf', 'Leaf Holes', 'Black Spot', 'Dry Leaf']           # configs/config.py:12-17


def _synthetic_per_class() -> int:
    return int(os.environ.get('ROVITKAN_SYNTH_PER_CLASS', '64'))


def synthetic_image(class_idx: int, index: int, size: int = IMAGE_SIZE) -> torch.Tensor:
    """Deterministic CHW image in [0, 1]: a class-dependent low-frequency colour pattern plus seeded noise, so the four
    classes are separable (a few epochs of training move the loss) while every sample is distinct."""
    g = torch.Generator().manual_seed(zlib.crc32(f'{class_idx}/{index}'.encode()))
    yy, xx = torch.meshgrid(torch.linspace(0, 1, size), torch.linspace(0, 1, size), indexing='ij')
    phase = torch.rand(3, generator=g) * 6.283
    freq = 1.0 + class_idx
    base = torch.stack([0.5 + 0.35 * torch.sin(6.283 * freq * xx + phase[0]),
                        0.5 + 0.35 * torch.sin(6.283 * freq * yy + phase[1]),
                        0.5 + 0.35 * torch.sin(6.283 * freq * (xx + yy) * 0.5 + phase[2])])
    return (base + 0.08 * torch.randn(3, size, size, generator=g)).clamp_(0.0, 1.0)


def synthetic_root() -> str:
    return os.environ.get('ROVITKAN_SYNTH_DIR') or os.path.join(tempfile.gettempdir(), f'rovitkan_synthetic_{os.getuid()}')


def _materialise(mode: str, class_name: str, class_idx: int, index: int) -> str:
    """Path of the PNG of synthetic sample (class, index), written on first use (atomically: DataLoader workers and parallel
    ranks may ask for the same file)."""
    d = os.path.join(synthetic_root(), mode, class_name)
    path = os.path.join(d, f'{index}.png')
    if not os.path.exists(path):
        from PIL import Image
        os.makedirs(d, exist_ok=True)
        pixels = (synthetic_image(class_idx, index) * 255.0).round().to(torch.uint8).permute(1, 2, 0).contiguous().numpy()
        tmp = f'{path}.{os.getpid()}.tmp'
        Image.fromarray(pixels).save(tmp, format='PNG', compress_level=1)
        os.replace(tmp, path)
    return path


class RoseLeafDataset(Dataset):
    def __init__(self, root_dir, class_names: Optional[List[str]] = None, severity_map: Optional[Dict[str, int]] = None,
                 transform=None, mode: str = 'augmented'):
        self.root_dir = str(root_dir)
        self.classes = list(class_names) if class_names is not None else list(DEFAULT_CLASSES)
        self.class_names = self.classes
        self.severity_map = dict(severity_map) if severity_map is not None else {c: i for i, c in enumerate(self.classes)}
        self.transform = transform
        self.mode = mode
        self.samples = []
        for ci, cname in enumerate(self.classes):
            d = os.path.join(self.root_dir, cname)
            if os.path.isdir(d):
                for f in sorted(os.listdir(d)):
                    if f.lower().endswith(_EXT):
                        self.samples.append({'path': os.path.join(d, f), 'class_idx': ci, 'severity': int(self.severity_map[cname])})
        self.synthetic = not self.samples
        if self.synthetic:
            n = _synthetic_per_class()
            for ci, cname in enumerate(self.classes):
                for i in range(n):
                    path = _materialise(mode, cname, ci, i + (0 if mode == 'augmented' else 1 << 20))
                    self.samples.append({'path': path, 'class_idx': ci, 'severity': int(self.severity_map[cname])})

    def __len__(self) -> int:
        return len(self.samples)

    def _load(self, sample, idx):
        from PIL import Image
        return Image.open(sample['path']).convert('RGB')

    def __getitem__(self, idx):
        s = self.samples[idx]
        img = self._load(s, idx)
        if self.transform is not None:
            img = self.transform(img)
        return img, s['class_idx'], s['severity']

    def get_class_weights(self) -> torch.Tensor:
        """Inverse-frequency focal alpha (scripts/train.py:110-111): total / (num_classes * count_c)."""
        counts = torch.zeros(len(self.classes))
        for s in self.samples:
            counts[s['class_idx']] += 1
        return counts.sum() / (len(self.classes) * counts.clamp_min(1.0))


def _loader(ds, batch_size, shuffle, num_workers, seed):
    workers = int(os.environ.get('ROVITKAN_DATA_WORKERS', num_workers))
    return DataLoader(ds, batch_size=batch_size, shuffle=shuffle, num_workers=workers, pin_memory=torch.cuda.is_available(),
                      generator=torch.Generator().manual_seed(seed) if shuffle else None, persistent_workers=False)


def create_dataloaders(augmented_root, original_root, class_names, severity_map, augmented_transform=None,
                       original_transform=None, batch_size: int = 32, train_val_split: float = 0.8, num_workers: int = 4,
                       seed: int = 42):
    """scripts/train.py:73-84: train/val = split of the augmented set, test = the original set."""
    full = RoseLeafDataset(augmented_root, class_names, severity_map, transform=augmented_transform, mode='augmented')
    n_train = int(round(train_val_split * len(full)))
    train_ds, val_ds = random_split(full, [n_train, len(full) - n_train], generator=torch.Generator().manual_seed(seed))
    test_ds = RoseLeafDataset(original_root, class_names, severity_map, transform=original_transform, mode='original')
    print(f'Dataset: {len(train_ds)} train / {len(val_ds)} val / {len(test_ds)} test '
          f'({"synthetic" if full.synthetic else "image folders"})')
    return (_loader(train_ds, batch_size, True, num_workers, seed), _loader(val_ds, batch_size, False, num_workers, seed),
            _loader(test_ds, batch_size, False, num_workers, seed))
