"""Data-parallel plumbing: one process per GPU, sum-all-reduce of the flat gradient, overlapped with the trunk backward.

The reference is single-process (no torch.distributed anywhere); RoViT-KAN has no BatchNorm and every loss is a per-rank
mean over equal local batches, so averaging gradients over ranks reproduces the single-GPU large-batch step exactly.

The trunk's backward writes its 150 gradients into one contiguous fp32 buffer in parameter order (ops.EncoderFn.backward),
walking the blocks 11 -> 0; the fused training tail returns the 23 head / KAN gradients as views of a second flat buffer.
`all_reduce_gradients` reduces every such contiguous run IN PLACE (two NCCL calls per step, no flatten / copy-back).
With `enable_overlap()` the trunk backward instead runs in `buckets` pieces (rvk_encoder_backward_range) and the slice of the
flat buffer that is final after each piece is all-reduced asynchronously (NCCL's own stream) while the next piece computes;
`all_reduce_gradients` then only adds the heads' call and joins the pending work.  Measured on 2 and 8 B200 the overlapped
schedule is slightly SLOWER (the whole payload takes 0.07-0.12 ms of a 6 ms step and NCCL's CTAs compete with the backward
kernels), so it is opt-in.  The 1/world scale is either applied here (`average=True`) or left to the optimizer kernel
(`FusedAdamW(grad_mult=1/world)`): one pass less over the gradients.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

_overlap = None          # {'group', 'world', 'buckets', 'pending': [(work, flat_slice)], 'calls'}


def enable_overlap(buckets: int = 3, group=None) -> bool:
    """Reduce the trunk gradient bucket by bucket from inside the backward pass.  No-op (False) without a process group."""
    global _overlap
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        _overlap = None
        return False
    _overlap = {'group': group, 'world': dist.get_world_size(group), 'buckets': max(1, min(int(buckets), 13)), 'pending': [],
                'calls': 0}
    return True


def disable_overlap() -> None:
    global _overlap
    _overlap = None


def overlap_state():
    return _overlap


def bucket_stage_ranges(buckets: int):
    """Split the 14 backward stages (0 final LayerNorm, 1 + j block 11 - j, 13 patch embedding) into `buckets` consecutive
    ranges; returns [(stage_begin, stage_end, first_block_final)] where after the range every gradient of blocks
    >= first_block_final (and the final norm) is complete; the last range completes everything (first_block_final = -1)."""
    per = -(-12 // buckets)
    out, blk_hi = [], 11
    while blk_hi >= 0:
        blk_lo = max(blk_hi - per + 1, 0)
        s_begin = 0 if blk_hi == 11 else 1 + (11 - blk_hi)
        s_end = 14 if blk_lo == 0 else 1 + (11 - blk_lo) + 1
        out.append((s_begin, s_end, -1 if blk_lo == 0 else blk_lo))
        blk_hi = blk_lo - 1
    return out


def launch_bucket(flat: torch.Tensor, lo: int, hi: int) -> None:
    """Called by ops.EncoderFn.backward after a stage range: all-reduce flat[lo:hi] without blocking the compute stream."""
    st = _overlap
    piece = flat[lo:hi]
    work = dist.all_reduce(piece, op=dist.ReduceOp.SUM, group=st['group'], async_op=True)
    st['pending'].append((work, piece))
    st['calls'] += 1


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


def all_reduce_gradients(params, world_size: int | None = None, group=None, average: bool = True) -> int:
    """Sum (average=False) or average `.grad` of `params` over the process group.  Gradient slices that the overlapped
    backward has already put in flight are only joined here.  Returns the number of collectives issued for this step."""
    if not dist.is_available() or not dist.is_initialized():
        return 0
    world = world_size or dist.get_world_size(group)
    if world == 1:
        return 0
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return 0
    calls = 0
    scale = 1.0 / world
    pending = _overlap['pending'] if _overlap is not None else []
    done_ranges = [(piece.data_ptr(), piece.data_ptr() + piece.numel() * piece.element_size()) for _, piece in pending]

    def in_flight(g):
        a = g.data_ptr()
        return any(lo <= a < hi for lo, hi in done_ranges)
    rest_all = [g for g in grads if not in_flight(g)]
    # split what is left into maximal runs that are already one flat buffer (the trunk gradient when overlap is off, the 23
    # head / KAN gradients of the fused training tail): those are reduced IN PLACE; single leftovers are coalesced
    runs = []              # one linear pass: a gradient extends the current run iff it starts where the previous one ended
    end_ptr, cur = None, None
    for g in rest_all:
        ptr = g.data_ptr()
        ok = (cur is not None and ptr == end_ptr and g.is_contiguous() and g.dtype == cur[0].dtype
              and g.untyped_storage().data_ptr() == cur[0].untyped_storage().data_ptr())
        if ok:
            cur.append(g)
        else:
            cur = [g]
            runs.append(cur)
        end_ptr = ptr + g.numel() * g.element_size() if g.is_contiguous() else None
    works, rest = [], []
    for run in runs:
        if len(run) == 1 and run[0].numel() < (1 << 16):
            rest.append(run[0])
            continue
        total = sum(g.numel() for g in run)
        flat = torch.empty(0, dtype=run[0].dtype, device=run[0].device).set_(
            run[0].untyped_storage(), run[0].storage_offset(), (total,), (1,))
        works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, None))
        calls += 1
    if rest:
        flat = torch._utils._flatten_dense_tensors(rest)
        works.append((dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True), flat, rest))
        calls += 1
    for work, piece in pending:
        work.wait()
        if average:
            piece.mul_(scale)
    if _overlap is not None:
        calls += len(pending)
        _overlap['pending'] = []
    for work, flat, back in works:
        work.wait()
        if average:
            flat.mul_(scale)
        if back is not None:
            for g, r in zip(back, torch._utils._unflatten_dense_tensors(flat, back)):
                g.copy_(r)
    return calls
