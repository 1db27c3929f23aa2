"""Host-side mirror of the reference's `models` package (same module and class names)."""

from .backbone import DeiTTinyBackbone, freeze_backbone, get_backbone_output_dim  # noqa: F401
from .heads import ClassificationHead, OrdinalHead, UncertaintyHead  # noqa: F401
from .kan import BSplineBasis, KANLayer, KANSeverityModule  # noqa: F401
from .rovit_kan import RoViTKAN  # noqa: F401
