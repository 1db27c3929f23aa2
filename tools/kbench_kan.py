#!/usr/bin/env python
"""Micro-benchmark of one KAN layer forward (and backward) at the microbenchmark batch.  usage: kbench_kan.py [batch] [in] [out] [bwd]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rovitkan_b200.models.kan import KANLayer

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n_in = int(sys.argv[2]) if len(sys.argv) > 2 else 192
n_out = int(sys.argv[3]) if len(sys.argv) > 3 else 64
bwd = len(sys.argv) > 4 and sys.argv[4] == 'bwd'
torch.manual_seed(0)
layer = KANLayer(n_in, n_out).cuda()
x = torch.randn(batch, n_in, device='cuda', requires_grad=bwd)
gy = torch.randn(batch, n_out, device='cuda')


def run():
    if bwd:
        y = layer(x)
        y.backward(gy)
        layer.zero_grad(set_to_none=True)
        x.grad = None
    else:
        with torch.no_grad():
            layer(x)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
print(f'KANLayer {n_in}->{n_out} batch {batch} {"fwd+bwd" if bwd else "fwd"}: {us:.1f} us  '
      f'{batch * n_in * 8 * n_out * 2 * (3 if bwd else 1) / us / 1e6:.1f} TFLOP/s dense-8  {batch * n_in * 4 / us / 1e3:.0f} GB/s of x')
