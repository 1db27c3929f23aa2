#!/usr/bin/env python
"""Micro-benchmark of the fused inference tail (rvk_heads_fused).  usage: kbench_heads.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rovitkan_b200 import ops
from rovitkan_b200.models import RoViTKAN

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
torch.manual_seed(0)
m = RoViTKAN(pretrained=False).cuda().eval()
f = torch.randn(batch, 192, device='cuda')
ps = m._fused_tail_params()
st = ops.HeadsFusedState()
kn = m.kan_module.kan_layers[0].knots_host()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3):
    ops.heads_fused(st, f, ps, kn)
torch.cuda.synchronize()
for cold in (False, True):
    tot = 0.0
    for _ in range(10):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.heads_fused(st, f, ps, kn)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    print(f'heads_fused batch {batch} ({"L2 flushed" if cold else "warm"}): {tot * 100:.1f} us per call (incl. output allocation + launch)')
