"""Oracle (CPU checker) for the composed RoViT-KAN forward.  TEST INFRASTRUCTURE ONLY.

Restates reference `models/rovit_kan.py:88-124` (`RoViTKAN.forward`, stage
gating) and `:126-161` (`predict`) on top of the other oracle modules, driven by
a flat `state_dict` that uses the reference's parameter names:

  backbone.model.*                         150 timm trunk tensors (oracle/vit.py)
  classification_head.{fc1,fc2}.{weight,bias}
  ordinal_head.{fc1,fc2}.{weight,bias}
  uncertainty_head.{fc1,fc_mu,fc_logvar}.{weight,bias}
  kan_module.kan_layers.{i}.{spline_weights,linear.weight,linear.bias,knots}
"""

from __future__ import annotations

import torch

from . import heads, kan, vit

KAN_DIMS = [192, 64, 16, 1]
HIDDEN = 128
NUM_CLASSES = 4


def random_state_dict(seed: int = 0, kan_dims=None, hidden: int = HIDDEN,
                      num_classes: int = NUM_CLASSES) -> dict:
    """Random-init weights with the reference's shapes, names and init laws."""
    kan_dims = list(kan_dims or KAN_DIMS)
    torch.manual_seed(seed)
    trunk = vit.DeiTTinyOracle()
    sd = {'backbone.model.' + k: v.detach().clone() for k, v in trunk.state_dict().items()}
    emb = vit.EMBED

    def lin(prefix, n_out, n_in):
        m = torch.nn.Linear(n_in, n_out)
        sd[prefix + '.weight'] = m.weight.detach().clone()
        sd[prefix + '.bias'] = m.bias.detach().clone()

    lin('classification_head.fc1', hidden, emb)
    lin('classification_head.fc2', num_classes, hidden)
    lin('ordinal_head.fc1', hidden, emb)
    lin('ordinal_head.fc2', num_classes - 1, hidden)
    lin('uncertainty_head.fc1', hidden, emb)
    lin('uncertainty_head.fc_mu', 1, hidden)
    lin('uncertainty_head.fc_logvar', 1, hidden)
    for i, (a, b) in enumerate(zip(kan_dims[:-1], kan_dims[1:])):
        p = f'kan_module.kan_layers.{i}.'
        sd[p + 'knots'] = kan.make_knots()
        sd[p + 'spline_weights'] = torch.randn(a, b, 7) * 0.1
        lin(p + 'linear', b, a)
    return sd


def kan_layers_from(sd: dict):
    layers, i = [], 0
    while f'kan_module.kan_layers.{i}.spline_weights' in sd:
        p = f'kan_module.kan_layers.{i}.'
        layers.append((sd[p + 'spline_weights'], sd[p + 'linear.weight'], sd[p + 'linear.bias']))
        i += 1
    return layers


def heads_forward(sd: dict, features: torch.Tensor, stage: int = 4, kan_loop: bool = False) -> dict:
    g = lambda k: sd[k]
    out = {'features': features}
    out['cls_logits'] = heads.classification_forward(
        features, g('classification_head.fc1.weight'), g('classification_head.fc1.bias'),
        g('classification_head.fc2.weight'), g('classification_head.fc2.bias'))
    out['ordinal_logits'] = None
    out['mu'] = out['log_var'] = None
    out['kan_severity'] = None
    if stage >= 2:
        out['ordinal_logits'] = heads.ordinal_forward(
            features, g('ordinal_head.fc1.weight'), g('ordinal_head.fc1.bias'),
            g('ordinal_head.fc2.weight'), g('ordinal_head.fc2.bias'))
    if stage >= 3:
        out['mu'], out['log_var'] = heads.uncertainty_forward(
            features, g('uncertainty_head.fc1.weight'), g('uncertainty_head.fc1.bias'),
            g('uncertainty_head.fc_mu.weight'), g('uncertainty_head.fc_mu.bias'),
            g('uncertainty_head.fc_logvar.weight'), g('uncertainty_head.fc_logvar.bias'))
    if stage >= 4:
        knots = sd['kan_module.kan_layers.0.knots']
        out['kan_severity'] = kan.severity_forward(features, kan_layers_from(sd), knots, loop=kan_loop)
    return out


def forward(sd: dict, images: torch.Tensor, stage: int = 4, kan_loop: bool = False) -> dict:
    """RoViTKAN.forward (rovit_kan.py:88-124), eval mode / dropout off."""
    features = vit.forward_functional(sd, images, prefix='backbone.model.')
    return heads_forward(sd, features, stage, kan_loop)


def predict(sd: dict, images: torch.Tensor, stage: int = 4) -> dict:
    """RoViTKAN.predict (rovit_kan.py:126-161)."""
    with torch.no_grad():
        o = forward(sd, images, stage)
        probs = torch.softmax(o['cls_logits'], dim=1)
        p = {'class': probs.argmax(dim=1), 'class_probs': probs, 'features': o['features']}
        if o['ordinal_logits'] is not None:
            p['ordinal_probs'] = heads.ordinal_probabilities(o['ordinal_logits'])
            p['ordinal_severity'] = heads.ordinal_severity(o['ordinal_logits'])
        if o['mu'] is not None:
            p['uncertainty_mu'] = o['mu']
            p['uncertainty_std'] = torch.exp(0.5 * o['log_var'])
        if o['kan_severity'] is not None:
            p['kan_severity'] = o['kan_severity']
        return p
