"""`data.transforms` surface used by the reference (scripts/train.py:13,81-82; scripts/evaluate.py:11;
scripts/run_ablation.py:12; training/trainer.py:9,86-92)."""

from __future__ import annotations

import math

import numpy as np
import torch

IMAGE_SIZE = 224
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)


def _to_tensor(img) -> torch.Tensor:
    """PIL image / HWC uint8 array / CHW tensor -> float CHW tensor in [0, 1] at 224x224."""
    if isinstance(img, torch.Tensor):
        t = img.float()
        if t.dim() == 3 and t.shape[0] not in (1, 3) and t.shape[-1] in (1, 3):
            t = t.permute(2, 0, 1)
        if t.max() > 1.5:
            t = t / 255.0
    else:
        a = np.asarray(img.convert('RGB') if hasattr(img, 'convert') else img)
        t = torch.from_numpy(np.array(a, copy=True)).permute(2, 0, 1).float() / 255.0
    if t.shape[-2:] != (IMAGE_SIZE, IMAGE_SIZE):
        t = torch.nn.functional.interpolate(t[None], size=(IMAGE_SIZE, IMAGE_SIZE), mode='bilinear', align_corners=False,
                                            antialias=True)[0]
    return t


class _Pipeline:
    """Picklable transform (DataLoader workers): resize -> [flips, brightness jitter] -> ImageNet normalisation."""

    def __init__(self, augment: bool):
        self.augment = augment

    def __call__(self, img) -> torch.Tensor:
        t = _to_tensor(img)
        if self.augment:
            if torch.rand(()) < 0.5:
                t = t.flip(-1)
            if torch.rand(()) < 0.5:
                t = t.flip(-2)
            t = (t * (0.8 + 0.4 * float(torch.rand(())))).clamp_(0.0, 1.0)
        mean = torch.tensor(MEAN).view(3, 1, 1)
        std = torch.tensor(STD).view(3, 1, 1)
        return (t - mean) / std

    def __repr__(self):
        return f'{type(self).__name__}(augment={self.augment})'


def augmented_transforms():
    return _Pipeline(augment=True)


def original_transforms():
    return _Pipeline(augment=False)


def inference_transforms():
    return _Pipeline(augment=False)


def _rand_box(h: int, w: int, lam: float):
    cut = math.sqrt(max(0.0, 1.0 - lam))
    ch, cw = int(h * cut), int(w * cut)
    cy, cx = int(np.random.randint(h)), int(np.random.randint(w))
    y0, y1 = max(cy - ch // 2, 0), min(cy + ch // 2, h)
    x0, x1 = max(cx - cw // 2, 0), min(cx + cw // 2, w)
    return y0, y1, x0, x1


def cutmix_or_mixup(images: torch.Tensor, labels: torch.Tensor, use_cutmix: bool = True, use_mixup: bool = True,
                    cutmix_alpha: float = 1.0, mixup_alpha: float = 0.2):
    """training/trainer.py:86-92 -> (mixed images, labels_a, labels_b, lam); lam is a Python float because the trainer
    blends the two loss dicts with it (trainer.py:107-110).  Runs on the tensors' own device (no host round trip)."""
    if not (use_cutmix or use_mixup) or images.shape[0] < 2:
        return images, labels, labels, 1.0
    pick_cutmix = use_cutmix and (not use_mixup or np.random.rand() < 0.5)
    perm = torch.randperm(images.shape[0], device=images.device)
    labels_b = labels[perm]
    if pick_cutmix:
        lam = float(np.random.beta(cutmix_alpha, cutmix_alpha)) if cutmix_alpha > 0 else 1.0
        h, w = images.shape[-2:]
        y0, y1, x0, x1 = _rand_box(h, w, lam)
        mixed = images.clone()
        mixed[..., y0:y1, x0:x1] = images[perm][..., y0:y1, x0:x1]
        lam = 1.0 - (y1 - y0) * (x1 - x0) / float(h * w)
    else:
        lam = float(np.random.beta(mixup_alpha, mixup_alpha)) if mixup_alpha > 0 else 1.0
        mixed = torch.lerp(images[perm], images, lam)
    return mixed, labels, labels_b, lam
