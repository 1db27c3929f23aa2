#!/usr/bin/env python
"""Micro-benchmark (and optional clock trace) of the attention forward kernel, or of the backward kernel.
usage: kbench_attn.py [images=1024] [bwd]   (RVK_ATTN_BWD_SIMT=1: the mma.sync backward)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rovitkan_b200 import _lib

images = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = images * 197
qkv = (torch.randn(M, 576, device='cuda') * 1.0).to(torch.bfloat16)
ctx = torch.empty(M, 192, device='cuda', dtype=torch.bfloat16)
s = torch.cuda.current_stream().cuda_stream
bwd = len(sys.argv) > 2 and sys.argv[2] == 'bwd'
if bwd:
    lse = torch.zeros(images, 3, 197, device='cuda')
    dctx = torch.randn(M, 192, device='cuda').to(torch.bfloat16)
    dqkv = torch.empty(M, 576, device='cuda', dtype=torch.bfloat16)
    _lib.call('rvk_attention_forward', qkv.data_ptr(), ctx.data_ptr(), lse.data_ptr(), images, s)


def run():
    if bwd:
        _lib.call('rvk_attention_backward', qkv.data_ptr(), ctx.data_ptr(), dctx.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), images, s)
    else:
        _lib.call('rvk_attention_forward', qkv.data_ptr(), ctx.data_ptr(), 0, images, s)


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
if bwd:
    print(f'attention bwd {images} images: {us:.1f} us/launch  {images * 3 * 10 * 197 * 197 * 64 / us / 1e6:.1f} TFLOP/s (5 algorithmic GEMMs)')
    sys.exit(0)
print(f'attention fwd {images} images: {us:.1f} us/launch  {images * 3 * 4 * 197 * 197 * 64 / us / 1e6:.1f} TFLOP/s')
if os.environ.get('ATTN_TRACE'):
    tr = torch.zeros(4 * 512, dtype=torch.int64, device='cuda')
    _lib.load().rvk_debug_set_attn_trace(tr.data_ptr())
    run()
    torch.cuda.synchronize()
    _lib.load().rvk_debug_set_attn_trace(0)
    t = tr.cpu().view(4, 512)
    for role, name in ((0, 'mma'), (1, 'smx')):
        ev = [(int(v) >> 48, int(v) & 0xFFFFFFFFFFFF) for v in t[role].tolist() if v != 0]
        if ev:
            t0 = ev[0][1]
            print(name, ' '.join(f'{tag}@{(c - t0)}' for tag, c in ev[:120]))
