"""Host-side mirror of the part of the reference's `training` package that is on the hot path (`losses`), plus the fused
optimizer tail and device-side step accounting (`optim`, `fast_trainer`: SURVEY.md N2 / N3)."""

from .losses import FocalLoss, JointLoss, KANRegressionLoss, OrdinalBCELoss, UncertaintyLoss  # noqa: F401
from .optim import FusedAdamW, StepStats  # noqa: F401
from .fast_trainer import FastTrainer  # noqa: F401
