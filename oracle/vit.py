"""Oracle (CPU checker) for the DeiT-Tiny trunk.  TEST INFRASTRUCTURE ONLY.

The reference does not contain this arithmetic: `models/backbone.py:12-16`
calls `timm.create_model('deit_tiny_patch16_224', pretrained, num_classes=0)`
and `requirements.txt:9` asks for `timm>=0.6.0` (un-pinned, un-vendored, not
installed in this image).  What follows restates the published architecture of
that timm model (timm `vision_transformer.py` / `deit.py`):

  PatchEmbed   Conv2d(3, 192, k=16, s=16, bias) -> flatten -> (B, 196, 192)
  _pos_embed   cat(cls_token (1,1,192), x); + pos_embed (1,197,192)
  Block x12    x = x + proj(attn(LN1(x)));  x = x + fc2(GELU_erf(fc1(LN2(x))))
               LayerNorm eps 1e-6, qkv bias, 3 heads of 64, scale 64**-0.5,
               no LayerScale, no DropPath, all dropouts p=0
  norm, pool   LN(x)[:, 0]  (global_pool='token', fc_norm/head = Identity
               because num_classes=0)

Parameter names are timm's, so a `state_dict` moves between this oracle, a real
timm model and the CUDA implementation unchanged.  Because timm itself is
unavailable the trunk is "parity unpinned" against timm; it is pinned against
torchvision's `VisionTransformer` and HuggingFace's `ViTModel` (both in the
image, both 5 524 416 parameters at these sizes) by `tests/test_oracle_golden.py`
through the weight remaps at the bottom of this file.
"""

from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

EMBED = 192
HEADS = 3
HEAD_DIM = 64
DEPTH = 12
MLP = 768
PATCH = 16
IMG = 224
TOKENS = 197
LN_EPS = 1e-6


class _Attention(nn.Module):
    def __init__(self, dim=EMBED, heads=HEADS):
        super().__init__()
        self.num_heads = heads
        self.head_dim = dim // heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.attn_drop = nn.Dropout(0.0)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x):
        b, n, c = x.shape
        qkv = self.qkv(x).reshape(b, n, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        att = (q * self.scale) @ k.transpose(-2, -1)
        att = att.softmax(dim=-1)
        x = (att @ v).transpose(1, 2).reshape(b, n, c)
        return self.proj(x)


class _Mlp(nn.Module):
    def __init__(self, dim=EMBED, hidden=MLP):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()          # exact erf form
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim=EMBED, heads=HEADS, hidden=MLP):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=LN_EPS)
        self.attn = _Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=LN_EPS)
        self.mlp = _Mlp(dim, hidden)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


class _PatchEmbed(nn.Module):
    def __init__(self, dim=EMBED):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=PATCH, stride=PATCH, bias=True)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class DeiTTinyOracle(nn.Module):
    """timm `deit_tiny_patch16_224` with `num_classes=0`, restated."""

    def __init__(self):
        super().__init__()
        self.num_features = EMBED
        self.embed_dim = EMBED
        self.patch_embed = _PatchEmbed()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, EMBED))
        self.pos_embed = nn.Parameter(torch.zeros(1, TOKENS, EMBED))
        self.blocks = nn.Sequential(*[_Block() for _ in range(DEPTH)])
        self.norm = nn.LayerNorm(EMBED, eps=LN_EPS)
        self._timm_init()

    def _timm_init(self):
        # timm init_weights(''): trunc_normal_(std=.02) on pos_embed and Linear
        # weights, zero biases, cls_token ~ N(0, 1e-6); Conv2d keeps torch's default.
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.normal_(self.cls_token, std=1e-6)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward_tokens(self, x):
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        x = x + self.pos_embed
        return self.blocks(x)

    def forward(self, x):
        return self.norm(self.forward_tokens(x))[:, 0]


def create_model(name: str, pretrained: bool = False, num_classes: int = 0, **_):
    """Stand-in for `timm.create_model` on the one model name the reference uses
    (models/backbone.py:12-16).  `pretrained` cannot download anything here."""
    if name != 'deit_tiny_patch16_224':
        raise ValueError(f'oracle timm shim only restates deit_tiny_patch16_224, got {name!r}')
    if num_classes != 0:
        raise ValueError('oracle timm shim only restates num_classes=0')
    return DeiTTinyOracle()


def forward_functional(sd: dict, images: torch.Tensor, prefix: str = '',
                       return_tokens: bool = False) -> torch.Tensor:
    """Same trunk from a flat `state_dict` (timm names, optional prefix), written
    as explicit tensor algebra.  Used to cross-check the module form and as the
    checker for individual CUDA kernels (per-stage outputs)."""
    g = lambda k: sd[prefix + k]
    b = images.shape[0]
    patches = images.reshape(b, 3, IMG // PATCH, PATCH, IMG // PATCH, PATCH)
    patches = patches.permute(0, 2, 4, 1, 3, 5).reshape(b, 196, 3 * PATCH * PATCH)
    x = patches @ g('patch_embed.proj.weight').reshape(EMBED, -1).t() + g('patch_embed.proj.bias')
    x = torch.cat([g('cls_token').expand(b, -1, -1), x], dim=1) + g('pos_embed')
    for i in range(DEPTH):
        p = f'blocks.{i}.'
        h = F.layer_norm(x, (EMBED,), g(p + 'norm1.weight'), g(p + 'norm1.bias'), LN_EPS)
        qkv = h @ g(p + 'attn.qkv.weight').t() + g(p + 'attn.qkv.bias')
        qkv = qkv.reshape(b, TOKENS, 3, HEADS, HEAD_DIM).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        att = torch.softmax((q @ k.transpose(-2, -1)) * (HEAD_DIM ** -0.5), dim=-1)
        ctx = (att @ v).transpose(1, 2).reshape(b, TOKENS, EMBED)
        x = x + ctx @ g(p + 'attn.proj.weight').t() + g(p + 'attn.proj.bias')
        h = F.layer_norm(x, (EMBED,), g(p + 'norm2.weight'), g(p + 'norm2.bias'), LN_EPS)
        z = h @ g(p + 'mlp.fc1.weight').t() + g(p + 'mlp.fc1.bias')
        a = 0.5 * z * (1.0 + torch.erf(z / math.sqrt(2.0)))
        x = x + a @ g(p + 'mlp.fc2.weight').t() + g(p + 'mlp.fc2.bias')
    if return_tokens:
        return x
    return F.layer_norm(x, (EMBED,), g('norm.weight'), g('norm.bias'), LN_EPS)[:, 0]


# ---------------------------------------------------------------------------
# Weight remaps onto the two independent implementations available offline.
# ---------------------------------------------------------------------------

def to_torchvision(sd: dict) -> dict:
    """timm-named trunk state_dict -> torchvision VisionTransformer names."""
    out = {
        'class_token': sd['cls_token'],
        'conv_proj.weight': sd['patch_embed.proj.weight'],
        'conv_proj.bias': sd['patch_embed.proj.bias'],
        'encoder.pos_embedding': sd['pos_embed'],
        'encoder.ln.weight': sd['norm.weight'],
        'encoder.ln.bias': sd['norm.bias'],
    }
    for i in range(DEPTH):
        s, d = f'blocks.{i}.', f'encoder.layers.encoder_layer_{i}.'
        out[d + 'ln_1.weight'] = sd[s + 'norm1.weight']
        out[d + 'ln_1.bias'] = sd[s + 'norm1.bias']
        out[d + 'self_attention.in_proj_weight'] = sd[s + 'attn.qkv.weight']
        out[d + 'self_attention.in_proj_bias'] = sd[s + 'attn.qkv.bias']
        out[d + 'self_attention.out_proj.weight'] = sd[s + 'attn.proj.weight']
        out[d + 'self_attention.out_proj.bias'] = sd[s + 'attn.proj.bias']
        out[d + 'ln_2.weight'] = sd[s + 'norm2.weight']
        out[d + 'ln_2.bias'] = sd[s + 'norm2.bias']
        out[d + 'mlp.0.weight'] = sd[s + 'mlp.fc1.weight']
        out[d + 'mlp.0.bias'] = sd[s + 'mlp.fc1.bias']
        out[d + 'mlp.3.weight'] = sd[s + 'mlp.fc2.weight']
        out[d + 'mlp.3.bias'] = sd[s + 'mlp.fc2.bias']
    return out


def to_hf_vit(sd: dict) -> dict:
    """timm-named trunk state_dict -> transformers.ViTModel names."""
    out = {
        'embeddings.cls_token': sd['cls_token'],
        'embeddings.position_embeddings': sd['pos_embed'],
        'embeddings.patch_embeddings.projection.weight': sd['patch_embed.proj.weight'],
        'embeddings.patch_embeddings.projection.bias': sd['patch_embed.proj.bias'],
        'layernorm.weight': sd['norm.weight'],
        'layernorm.bias': sd['norm.bias'],
    }
    for i in range(DEPTH):
        s, d = f'blocks.{i}.', f'encoder.layer.{i}.'
        wq, wk, wv = sd[s + 'attn.qkv.weight'].chunk(3, dim=0)
        bq, bk, bv = sd[s + 'attn.qkv.bias'].chunk(3, dim=0)
        out[d + 'attention.attention.query.weight'] = wq
        out[d + 'attention.attention.query.bias'] = bq
        out[d + 'attention.attention.key.weight'] = wk
        out[d + 'attention.attention.key.bias'] = bk
        out[d + 'attention.attention.value.weight'] = wv
        out[d + 'attention.attention.value.bias'] = bv
        out[d + 'attention.output.dense.weight'] = sd[s + 'attn.proj.weight']
        out[d + 'attention.output.dense.bias'] = sd[s + 'attn.proj.bias']
        out[d + 'layernorm_before.weight'] = sd[s + 'norm1.weight']
        out[d + 'layernorm_before.bias'] = sd[s + 'norm1.bias']
        out[d + 'layernorm_after.weight'] = sd[s + 'norm2.weight']
        out[d + 'layernorm_after.bias'] = sd[s + 'norm2.bias']
        out[d + 'intermediate.dense.weight'] = sd[s + 'mlp.fc1.weight']
        out[d + 'intermediate.dense.bias'] = sd[s + 'mlp.fc1.bias']
        out[d + 'output.dense.weight'] = sd[s + 'mlp.fc2.weight']
        out[d + 'output.dense.bias'] = sd[s + 'mlp.fc2.bias']
    return out
