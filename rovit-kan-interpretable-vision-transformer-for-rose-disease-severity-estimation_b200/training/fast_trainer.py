"""Sync-free variant of the reference's epoch loop (reference training/trainer.py:20-340).

Same constructor, same `fit()` / `train_epoch()` / `val_epoch()` / `save_checkpoint()` / `load_checkpoint()` contract, same
metric keys, curriculum / freeze schedule, early stopping and checkpoint dictionary as the reference's `Trainer`, so it can be
swapped in where a script builds `Trainer(...)`.  What differs is only WHERE the host waits for the device:

  * the reference reads six scalars back per step (`loss.item()` x5 and the correct count, trainer.py:144-153): every one is a
    full stream synchronisation; here they are accumulated on the device (`StepStats`) and read once per epoch;
  * with `FusedAdamW(max_grad_norm=...)` the tail `scaler.unscale_ -> clip_grad_norm_ -> scaler.step` (trainer.py:122-128) is
    one fused pass (unscale + inf check + global-norm clip + AdamW) with no host round trip; any other optimizer is driven
    exactly as the reference drives it;
  * host->device copies are non-blocking, and under torch.distributed the flat trunk / head gradients are all-reduced in place
    (two NCCL calls, the 1/world scale folded into the optimizer kernel) -- the reference is single-process.
"""

from __future__ import annotations

from pathlib import Path
from typing import Dict

import torch
import torch.nn as nn
from torch.amp import GradScaler, autocast

from .. import dist as rdist
from ..data.transforms import cutmix_or_mixup
from .optim import FusedAdamW, StepStats


class FastTrainer:
    def __init__(self, model: nn.Module, train_loader, val_loader, optimizer, scheduler, loss_fn, config, device: torch.device,
                 logger=None):
        self.model = model.to(device)
        self.train_loader, self.val_loader = train_loader, val_loader
        self.optimizer, self.scheduler, self.loss_fn = optimizer, scheduler, loss_fn
        self.config, self.device, self.logger = config, device, logger
        # trainer.py:44-47: mixed precision only on CUDA
        self.scaler = GradScaler('cuda') if (config.flags.mixed_precision and device.type == 'cuda') else None
        self.best_val_loss = float('inf')
        self.patience_counter = 0
        self.best_epoch = 0
        self.world = torch.distributed.get_world_size() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        if self.world > 1 and isinstance(optimizer, FusedAdamW):
            optimizer.grad_mult = 1.0 / self.world

    # ------------------------------------------------------------------------------------------------ one epoch
    def _fused_tail(self) -> bool:
        return isinstance(self.optimizer, FusedAdamW) and bool(self.optimizer.max_grad_norm)

    def train_epoch(self, epoch: int) -> Dict[str, float]:
        cfg = self.config
        self.model.train()
        stage = cfg.get_stage_for_epoch(epoch)                                   # trainer.py:58-59
        self.model.curriculum_stage = stage
        if epoch == cfg.flags.freeze_backbone_epochs + 1:                        # trainer.py:62-63
            self.model.unfreeze_backbone()
        stats = StepStats(self.device)
        params = list(self.model.parameters())
        for images, class_labels, severity_labels in self.train_loader:
            images = images.to(self.device, non_blocking=True)
            class_labels = class_labels.to(self.device, non_blocking=True)
            severity_labels = severity_labels.to(self.device, non_blocking=True)
            mixed = cfg.flags.use_cutmix or cfg.flags.use_mixup                   # trainer.py:85-96
            if mixed:
                images, labels_a, labels_b, lam = cutmix_or_mixup(
                    images, class_labels, use_cutmix=cfg.flags.use_cutmix, use_mixup=cfg.flags.use_mixup,
                    cutmix_alpha=cfg.flags.cutmix_alpha, mixup_alpha=cfg.flags.mixup_alpha)
            with autocast('cuda', enabled=self.scaler is not None):
                outputs = self.model(images)
                if mixed:                                                        # trainer.py:104-111
                    la = self.loss_fn(outputs, labels_a, severity_labels, stage)
                    lb = self.loss_fn(outputs, labels_b, severity_labels, stage)
                    losses = {k: lam * la[k] + (1 - lam) * lb[k] for k in la}
                else:
                    losses = self.loss_fn(outputs, class_labels, severity_labels, stage)
                loss = losses['total_loss']
            self.optimizer.zero_grad(set_to_none=True)
            (self.scaler.scale(loss) if self.scaler is not None else loss).backward()
            if self.world > 1:
                rdist.all_reduce_gradients(params, self.world, average=not isinstance(self.optimizer, FusedAdamW))
            if self._fused_tail():                                               # unscale + inf check + clip + AdamW: one pass
                if self.scaler is not None:
                    self.scaler.step(self.optimizer)
                    self.scaler.update()
                else:
                    self.optimizer.step()
            else:                                                                # trainer.py:122-129 / 137-141
                if self.scaler is not None:
                    self.scaler.unscale_(self.optimizer)
                torch.nn.utils.clip_grad_norm_(self.model.parameters(), cfg.flags.gradient_clip)
                if self.scaler is not None:
                    self.scaler.step(self.optimizer)
                    self.scaler.update()
                else:
                    self.optimizer.step()
            stats.update(losses, outputs['cls_logits'], class_labels)            # device-side adds only
        return stats.result()                                                    # the single device -> host read of the epoch

    def val_epoch(self) -> Dict[str, float]:
        self.model.eval()
        stats = StepStats(self.device)
        with torch.no_grad():
            for images, class_labels, severity_labels in self.val_loader:
                images = images.to(self.device, non_blocking=True)
                class_labels = class_labels.to(self.device, non_blocking=True)
                severity_labels = severity_labels.to(self.device, non_blocking=True)
                outputs = self.model(images)
                losses = self.loss_fn(outputs, class_labels, severity_labels, stage=4)     # trainer.py:205
                stats.update(losses, outputs['cls_logits'], class_labels)
        return stats.result()

    # ------------------------------------------------------------------------------------------------ the run
    def fit(self) -> Dict[str, list]:
        cfg = self.config
        if cfg.flags.freeze_backbone_epochs > 0:                                 # trainer.py:244-246
            self.model.freeze_backbone()
        history = {'train_loss': [], 'val_loss': [], 'train_acc': [], 'val_acc': []}
        for epoch in range(1, cfg.train.epochs + 1):
            train_metrics = self.train_epoch(epoch)
            val_metrics = self.val_epoch()
            self.scheduler.step()
            if self.logger:
                self.logger.log_epoch(epoch, cfg.get_stage_for_epoch(epoch), train_metrics, val_metrics)
            print(f"Epoch {epoch}/{cfg.train.epochs}  train loss {train_metrics['loss']:.4f} acc {train_metrics['accuracy']:.2f}%  "
                  f"val loss {val_metrics['loss']:.4f} acc {val_metrics['accuracy']:.2f}%")
            history['train_loss'].append(train_metrics['loss'])
            history['val_loss'].append(val_metrics['loss'])
            history['train_acc'].append(train_metrics['accuracy'])
            history['val_acc'].append(val_metrics['accuracy'])
            if val_metrics['loss'] < self.best_val_loss:                         # trainer.py:282-293
                self.best_val_loss = val_metrics['loss']
                self.patience_counter = 0
                self.best_epoch = epoch
                self.save_checkpoint(Path(cfg.paths.checkpoints_dir) / 'best_model.pth', epoch, val_metrics)
            else:
                self.patience_counter += 1
            if self.patience_counter >= cfg.train.early_stop_patience:
                print(f'Early stopping at epoch {epoch} (best epoch {self.best_epoch}, val loss {self.best_val_loss:.4f})')
                break
        return history

    def save_checkpoint(self, path: Path, epoch: int, metrics: Dict):           # trainer.py:311-325 (same keys)
        checkpoint = {'epoch': epoch, 'model_state_dict': self.model.state_dict(),
                      'optimizer_state_dict': self.optimizer.state_dict(), 'scheduler_state_dict': self.scheduler.state_dict(),
                      'best_val_loss': self.best_val_loss, 'metrics': metrics, 'config': self.config}
        if self.scaler is not None:
            checkpoint['scaler_state_dict'] = self.scaler.state_dict()
        torch.save(checkpoint, path)

    def load_checkpoint(self, path: Path):                                       # trainer.py:327-340
        checkpoint = torch.load(path, map_location=self.device, weights_only=False)
        self.model.load_state_dict(checkpoint['model_state_dict'])
        self.optimizer.load_state_dict(checkpoint['optimizer_state_dict'])
        self.scheduler.load_state_dict(checkpoint['scheduler_state_dict'])
        self.best_val_loss = checkpoint['best_val_loss']
        if self.scaler is not None and 'scaler_state_dict' in checkpoint:
            self.scaler.load_state_dict(checkpoint['scaler_state_dict'])
        return checkpoint
