"""Parameter container + fused forward for timm's `deit_tiny_patch16_224` (num_classes=0).

The reference gets this trunk from `timm.create_model` (models/backbone.py:12-16).  timm is a
third-party dependency that is not part of the reference tree; this module keeps timm's module tree
and parameter names (`cls_token`, `pos_embed`, `patch_embed.proj`, `blocks.{i}.{norm1,attn.qkv,
attn.proj,norm2,mlp.fc1,mlp.fc2}`, `norm`) so timm / reference checkpoints load unchanged, and runs
the whole trunk as one launch sequence of sm_100a kernels (csrc/encoder.cu).

The submodules are parameter holders: the fused path never calls `blocks[i].forward`, so forward
hooks on `blocks[i].attn` / `.norm1` (reference explainability code) do not fire -- documented
limitation, see DESIGN.md.
"""

import warnings

import torch
import torch.nn as nn

from .._bootstrap import ops as _ops

EMBED, HEADS, DEPTH, MLP, PATCH, TOKENS = 192, 3, 12, 768, 16, 197


class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError(f'{type(self).__name__} is a parameter holder of the fused DeiT-Tiny trunk; '
                           'call the trunk (DeiTTinyBackbone / VisionTransformerB200) instead.')


class Attention(_Holder):
    def __init__(self):
        super().__init__()
        self.num_heads = HEADS
        self.head_dim = EMBED // HEADS
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(EMBED, EMBED * 3, bias=True)
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(EMBED, EMBED)
        self.proj_drop = nn.Dropout(0.0)


class Mlp(_Holder):
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(EMBED, MLP)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(MLP, EMBED)


class Block(_Holder):
    def __init__(self):
        super().__init__()
        self.norm1 = nn.LayerNorm(EMBED, eps=1e-6)
        self.attn = Attention()
        self.norm2 = nn.LayerNorm(EMBED, eps=1e-6)
        self.mlp = Mlp()


class PatchEmbed(_Holder):
    def __init__(self):
        super().__init__()
        self.proj = nn.Conv2d(3, EMBED, kernel_size=PATCH, stride=PATCH, bias=True)


_BLOCK_PARAMS = ['norm1.weight', 'norm1.bias', 'attn.qkv.weight', 'attn.qkv.bias', 'attn.proj.weight',
                 'attn.proj.bias', 'norm2.weight', 'norm2.bias', 'mlp.fc1.weight', 'mlp.fc1.bias',
                 'mlp.fc2.weight', 'mlp.fc2.bias']


def param_names():
    """The 150 trunk tensors in the order the C ABI expects (== timm's state_dict order)."""
    names = ['cls_token', 'pos_embed', 'patch_embed.proj.weight', 'patch_embed.proj.bias']
    for i in range(DEPTH):
        names += [f'blocks.{i}.{n}' for n in _BLOCK_PARAMS]
    return names + ['norm.weight', 'norm.bias']


class VisionTransformerB200(nn.Module):
    def __init__(self):
        super().__init__()
        self.num_features = EMBED
        self.embed_dim = EMBED
        self.num_classes = 0
        self.patch_embed = PatchEmbed()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, EMBED))
        self.pos_embed = nn.Parameter(torch.zeros(1, TOKENS, EMBED))
        self.blocks = nn.Sequential(*[Block() for _ in range(DEPTH)])
        self.norm = nn.LayerNorm(EMBED, eps=1e-6)
        # timm init: trunc_normal(.02) on pos_embed / Linear weights, zero biases, cls ~ N(0, 1e-6)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)
        self._state = None
        self._names = param_names()

    def _ordered_params(self):
        named = dict(self.named_parameters())
        return [named[n] for n in self._names]

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        ops = _ops()
        if self._state is None:
            self._state = ops.EncoderState()
        if self._hooked_modules():
            return self._forward_with_hooks(x)
        return ops.EncoderFn.apply(self._state, *ops.nograd(x, *self._ordered_params()))

    # ---- hook mode -----------------------------------------------------------------------------
    def _hooked_modules(self):
        mods = [self.patch_embed, self.norm]
        for b in self.blocks:
            mods += [b, b.norm1, b.attn, b.norm2, b.mlp]
        return [m for m in mods if m._forward_hooks]

    def trace(self, x: torch.Tensor):
        """Forward that keeps every block's intermediate tensors: returns (features, record) with
        record[i] = {'input', 'norm1', 'attn_out', 'attn_probs', 'mid', 'norm2', 'mlp_out', 'output'} for block i and
        record['final'] = the token stream entering the last LayerNorm.  No autograd graph is built."""
        ops = _ops()
        if self._state is None:
            self._state = ops.EncoderState()
        with torch.no_grad():
            feats, saved, ws = ops.encoder_traced_forward(self._state, x, self._ordered_params())
            rec = {}
            for i in range(DEPTH):
                x_in, x_mid = saved(i, 0), saved(i, 4)
                x_out = saved(i + 1, 0)
                rec[i] = {'input': x_in, 'norm1': saved(i, 1).float(), 'attn_out': x_mid - x_in, 'qkv': saved(i, 2), 'mid': x_mid,
                          'norm2': saved(i, 5).float(), 'mlp_out': x_out - x_mid, 'output': x_out}
            rec['final'] = saved(DEPTH, 0)
        return feats, rec

    def _forward_with_hooks(self, x: torch.Tensor) -> torch.Tensor:
        """Forward hooks registered on `blocks[i]`, `.norm1`, `.attn`, `.norm2`, `.mlp` or `norm` (the reference's
        explainability code: backbone.py:51-53, attention_maps.py:31-33, gradcam.py:40) fire with the tensors the fused
        kernels produced for that module.  The tensors are detached (no autograd through them: backward hooks do not fire)
        and a hook's return value cannot replace the module output."""
        feats, rec = self.trace(x)

        def fire(mod, inp, out):
            for hook in list(mod._forward_hooks.values()):
                if hook(mod, (inp,), out) is not None:
                    raise RuntimeError('forward hooks of the fused DeiT-Tiny trunk cannot replace the module output')
        for i, b in enumerate(self.blocks):
            r = rec[i]
            fire(b.norm1, r['input'], r['norm1'])
            fire(b.attn, r['norm1'], r['attn_out'])
            fire(b.norm2, r['mid'], r['norm2'])
            fire(b.mlp, r['norm2'], r['mlp_out'])
            fire(b, r['input'], r['output'])
        fire(self.norm, rec['final'], feats)
        return feats

    def attention_probabilities(self, x: torch.Tensor):
        """softmax(q k^T / sqrt(64)) of every block, each (B, 3, 197, 197) fp32: what timm's Attention calls `attn`."""
        ops = _ops()
        _, rec = self.trace(x)
        return [ops.attention_probabilities(rec[i]['qkv']) for i in range(DEPTH)]


PRETRAINED_ENV = 'ROVITKAN_PRETRAINED'


def _pretrained_state_dict(name: str):
    """Weights for `pretrained=True` (the reference's default, configs/config.py:60), without a network:
    $ROVITKAN_PRETRAINED = path of a timm `deit_tiny_patch16_224` checkpoint (a state_dict, or a dict holding one under
    'model' / 'state_dict' / 'model_state_dict'; classifier `head.*` keys are ignored), or `timm` itself if it is
    importable (its own cache / download), or the literal `random` to opt in to a randomly initialised trunk."""
    import os
    src = os.environ.get(PRETRAINED_ENV, '')
    if src.lower() == 'random':
        warnings.warn(f'{PRETRAINED_ENV}=random: pretrained=True was requested but the trunk is randomly initialised', stacklevel=4)
        return None
    if src:
        obj = torch.load(src, map_location='cpu', weights_only=False)
        for key in ('model', 'state_dict', 'model_state_dict'):
            if isinstance(obj, dict) and key in obj and isinstance(obj[key], dict):
                obj = obj[key]
        sd = {k[len('backbone.model.'):] if k.startswith('backbone.model.') else k: v for k, v in obj.items()}
        return {k: v for k, v in sd.items() if not k.startswith(('head.', 'head_dist.', 'fc_norm.'))}
    try:
        import timm
    except ImportError:
        raise RuntimeError(
            f"pretrained=True needs ImageNet weights for {name}, and this machine has neither `timm` nor a checkpoint: set "
            f"{PRETRAINED_ENV}=/path/to/deit_tiny_patch16_224.pth (timm parameter names), or {PRETRAINED_ENV}=random to train "
            f"from a random initialisation on purpose, or construct the model with pretrained=False.") from None
    return timm.create_model(name, pretrained=True, num_classes=0).state_dict()


def create_model(name: str = 'deit_tiny_patch16_224', pretrained: bool = False, num_classes: int = 0, **kwargs):
    """Stand-in for the one `timm.create_model` call the reference makes (models/backbone.py:12-16)."""
    if name != 'deit_tiny_patch16_224' or num_classes != 0:
        raise NotImplementedError('only deit_tiny_patch16_224 with num_classes=0 (what RoViT-KAN uses) is built')
    model = VisionTransformerB200()
    if pretrained:
        sd = _pretrained_state_dict(name)
        if sd is not None:
            model.load_state_dict(sd, strict=True)
    return model
