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
        """backbone.py:36-62: the outputs of `blocks[i].attn` captured by forward hooks -- with timm's Attention that is the
        attention block output (B, 197, 192) per block (timm returns no weights).  Same mechanism here: the hooks fire from
        the trunk's hook mode.  For the attention PROBABILITIES (B, 3, 197, 197) use `self.model.attention_probabilities(x)`."""
        attention_maps = []

        def hook_fn(module, input, output):
            attention_maps.append(output[1] if isinstance(output, tuple) else output)
        hooks = [block.attn.register_forward_hook(hook_fn) for block in self.model.blocks]
        try:
            _ = self.model(x)
        finally:
            for hook in hooks:
                hook.remove()
        return attention_maps


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
