"""Host-side mirror of the part of the reference's `training` package that is on the hot path."""

from .losses import FocalLoss, JointLoss, KANRegressionLoss, OrdinalBCELoss, UncertaintyLoss  # noqa: F401
