"""DeiT-Tiny backbone wrapper -- mirror of reference `models/backbone.py` (backbone.py:7-82).

`self.model` has timm's module tree and parameter names, so state_dicts keep the reference's
`backbone.model.*` keys; `forward` runs the fused sm_100a trunk (models/deit.py -> csrc/encoder.cu).
"""

import torch
import torch.nn as nn

from . import deit


class DeiTTinyBackbone(nn.Module):
    def __init__(self, pretrained: bool = True, freeze: bool = False):
        super().__init__()
        self.model = deit.create_model('deit_tiny_patch16_224', pretrained=pretrained, num_classes=0)
        self.embed_dim = self.model.num_features
        if freeze:
            self.freeze()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.model(x)

    def freeze(self):
        for param in self.model.parameters():
            param.requires_grad = False
        print("Backbone frozen")

    def unfreeze(self):
        for param in self.model.parameters():
            param.requires_grad = True
        print("Backbone unfrozen")

    def get_attention_maps(self, x: torch.Tensor):
        raise NotImplementedError(
            'attention-map capture relies on forward hooks on blocks[i].attn (backbone.py:51-53); the fused '
            'trunk keeps attention probabilities on chip and never materialises them. Out of scope for the '
            'forward/backward hot path (see DESIGN.md).')


def freeze_backbone(model: nn.Module, freeze: bool = True):
    if not hasattr(model, 'backbone'):
        raise AttributeError("Model does not have 'backbone' attribute")
    if freeze:
        model.backbone.freeze()
    else:
        model.backbone.unfreeze()


def get_backbone_output_dim(backbone_name: str = 'deit_tiny_patch16_224') -> int:
    # values as published by the reference (backbone.py:74-80), including its 384 entry for deit_tiny
    return {'deit_tiny_patch16_224': 384, 'deit_small_patch16_224': 384, 'deit_base_patch16_224': 768}.get(
        backbone_name, 384)
