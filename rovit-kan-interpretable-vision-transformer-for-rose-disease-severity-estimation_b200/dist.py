"""Data-parallel plumbing: one process per GPU, one sum-all-reduce of the flat gradient per step.

The reference is single-process (no torch.distributed anywhere); RoViT-KAN has no BatchNorm and every
loss is a per-rank mean over equal local batches, so averaging gradients over ranks reproduces the
single-GPU large-batch step exactly.  The trunk's backward already writes its 150 gradients into one
contiguous fp32 buffer (ops.EncoderFn.backward); when autograd hands those views to `.grad` unchanged
they are reduced in place with a single NCCL call, everything else is coalesced into one flat buffer.
"""

from __future__ import annotations

import torch
import torch.distributed as dist


def _contiguous_run(grads):
    """True if `grads` are back-to-back views of one allocation (the trunk's flat gradient)."""
    if not grads:
        return False
    base = grads[0].data_ptr()
    off = 0
    for g in grads:
        if g.data_ptr() != base + off or not g.is_contiguous() or g.dtype != grads[0].dtype:
            return False
        off += g.numel() * g.element_size()
    st = grads[0].untyped_storage()
    return base >= st.data_ptr() and base + off <= st.data_ptr() + st.nbytes()


def all_reduce_gradients(params, world_size: int | None = None, group=None) -> int:
    """Average `.grad` of `params` over the process group.  Returns the number of collectives issued."""
    if not dist.is_available() or not dist.is_initialized():
        return 0
    world = world_size or dist.get_world_size(group)
    if world == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    calls = 0
    # longest prefix that is already one flat buffer
    n = len(grads)
    while n > 0 and not _contiguous_run(grads[:n]):
        n -= 1 if n <= 150 else n - 150
    if n > 1:
        total = sum(g.numel() for g in grads[:n])
        flat = torch.empty(0, dtype=grads[0].dtype, device=grads[0].device).set_(
            grads[0].untyped_storage(), grads[0].storage_offset(), (total,), (1,))
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / world)
        calls += 1
    else:
        n = 0
    rest = grads[n:]
    if rest:
        flat = torch._utils._flatten_dense_tensors(rest)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / world)
        for g, r in zip(rest, torch._utils._unflatten_dense_tensors(flat, rest)):
            g.copy_(r)
        calls += 1
    return calls
