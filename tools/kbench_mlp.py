#!/usr/bin/env python
"""Micro-benchmark of the fused MLP block kernel alone (rvk_mlp_fused) at the inference batch's token count.
usage: python tools/kbench_mlp.py [cta_group=2] [images=1024] [iters=10] [proj]   (MLP_TRACE=1: clock64 event log of CTA 0)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rovitkan_b200 import _lib

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
images = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
proj = len(sys.argv) > 4 and sys.argv[4] == 'proj'
M = images * 197
dev = 'cuda'
g = torch.Generator().manual_seed(0)
w1 = (torch.randn(768, 192, generator=g) * 0.08).to(dev).to(torch.bfloat16)
w2 = (torch.randn(192, 768, generator=g) * 0.05).to(dev).to(torch.float16)
b1 = torch.zeros(768, device=dev)
b2 = torch.zeros(192, device=dev)
gamma = torch.ones(192, device=dev)
beta = torch.zeros(192, device=dev)
Mp = (M + 127) // 128 * 128
x = torch.randn(Mp * 192, device=dev)
ln_out = torch.empty(M, 192, device=dev, dtype=torch.bfloat16)
ctx = torch.randn(M, 192, generator=g).to(dev).to(torch.bfloat16)
wp = (torch.randn(192, 192, generator=g) * 0.02).to(dev).to(torch.bfloat16)
s = torch.cuda.current_stream().cuda_stream


def run():
    if proj:
        _lib.call('rvk_attn_proj_mlp_fused', x.data_ptr(), x.data_ptr(), ctx.data_ptr(), wp.data_ptr(), b2.data_ptr(),
                  gamma.data_ptr(), beta.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                  gamma.data_ptr(), beta.data_ptr(), 1e-6, ln_out.data_ptr(), M, G, s)
        return
    _lib.call('rvk_mlp_fused', x.data_ptr(), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), w1.data_ptr(), b1.data_ptr(),
              w2.data_ptr(), b2.data_ptr(), gamma.data_ptr(), beta.data_ptr(), 1e-6, ln_out.data_ptr(), M, G, s)


for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    run()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / iters
flops = 2.0 * M * 192 * (768 * 2 + (192 if proj else 0))
print(f'mlp_fused G={G} M={M} proj={proj}: {us:.1f} us/launch  {flops / us / 1e6:.1f} TFLOP/s  '
      f'{(M * 192 * (4 + 4 + 2)) / us / 1e3:.0f} GB/s algorithmic')

if os.environ.get('MLP_TRACE'):
    tr = torch.zeros(4 * 512, dtype=torch.int64, device=dev)
    _lib.load().rvk_debug_set_mlp_trace(tr.data_ptr())
    run()
    torch.cuda.synchronize()
    _lib.load().rvk_debug_set_mlp_trace(0)
    t = tr.cpu().view(4, 512)
    for role, name in ((0, 'mma'), (1, 'epi'), (2, 'epi1')):
        ev = [(int(v) >> 48, int(v) & 0xFFFFFFFFFFFF) for v in t[role].tolist() if v != 0]
        if not ev:
            continue
        t0 = ev[0][1]
        print(name, ' '.join(f'{tag}@{(c - t0)}' for tag, c in ev[:400]))
