"""world_size-2 gloo test (CPU) of the data-parallel gradient all-reduce helper."""

import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from rovitkan_b200.dist import _contiguous_run, all_reduce_gradients
    torch.manual_seed(0)
    # 150 "trunk" params whose grads are views of one flat buffer + 5 separately allocated "head" params
    shapes = [(3, 4), (7,), (2, 5, 2)] * 50
    params = [torch.nn.Parameter(torch.zeros(s)) for s in shapes] + [torch.nn.Parameter(torch.zeros(6, 3)) for _ in range(5)]
    flat = torch.zeros(sum(p.numel() for p in params[:150]))
    off = 0
    for p in params[:150]:
        p.grad = flat[off:off + p.numel()].view(p.shape)
        off += p.numel()
    for p in params[150:]:
        p.grad = torch.zeros_like(p)
    for i, p in enumerate(params):
        p.grad.fill_(float((rank + 1) * (i + 1)))
    assert _contiguous_run([p.grad for p in params[:150]])
    calls = all_reduce_gradients(params, world)
    want = [(1 + 2) / 2 * (i + 1) for i in range(len(params))]
    ok = all(torch.allclose(p.grad, torch.full_like(p.grad, w)) for p, w in zip(params, want))
    # ---- overlapped mode: the trunk slices are put in flight bucket by bucket (as ops.EncoderFn.backward does after each
    # stage range), all_reduce_gradients then only adds the head reduce and joins; average=False leaves the 1/world scale to
    # the optimizer kernel (FusedAdamW(grad_mult=1/world))
    from rovitkan_b200 import dist as rdist
    assert rdist.enable_overlap(buckets=3)
    assert rdist.bucket_stage_ranges(3) == [(0, 5, 8), (5, 9, 4), (9, 14, -1)]
    for i, p in enumerate(params):
        p.grad.fill_(float((rank + 1) * (i + 1)))
    cuts = [len(flat), 600, 250, 0]
    for hi, lo in zip(cuts[:-1], cuts[1:]):
        rdist.launch_bucket(flat, lo, hi)
    calls2 = all_reduce_gradients(params, world, average=False)
    ok2 = all(torch.allclose(p.grad, torch.full_like(p.grad, 2 * w)) for p, w in zip(params, want))
    ok2 = ok2 and rdist.overlap_state()['pending'] == []
    rdist.disable_overlap()
    q.put((rank, ok and ok2, (calls, calls2)))
    dist.destroy_process_group()


def test_all_reduce_gradients_world2_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    # plain: one flat trunk reduce + one coalesced heads reduce; overlapped: three trunk buckets + the heads reduce
    assert all(calls == (2, 4) for _, _, calls in res), res


def test_bucket_stage_ranges_tile_the_backward_for_every_bucket_count():
    """The overlapped schedule cuts the 14 backward stages (0 = final LayerNorm, 1 + j = block 11 - j, 13 = patch embedding)
    into consecutive ranges; whatever the bucket count, the ranges must tile [0, 14) in order, every cut must fall on a block
    boundary, and the block index from which the flat gradient is final must fall strictly until the last range (-1 = all)."""
    sys.path.insert(0, ROOT)
    from rovitkan_b200.dist import bucket_stage_ranges
    for buckets in range(1, 14):
        ranges = bucket_stage_ranges(buckets)
        assert 1 <= len(ranges) <= min(buckets, 12)
        assert ranges[0][0] == 0 and ranges[-1][1] == 14 and ranges[-1][2] == -1
        for (b0, e0, f0), (b1, e1, f1) in zip(ranges[:-1], ranges[1:]):
            assert e0 == b1 and b0 < e0 and b1 < e1
            assert f0 > f1 and f0 == 11 - (e0 - 2)          # stage e0 - 1 was block f0: blocks >= f0 are complete
    assert bucket_stage_ranges(1) == [(0, 14, -1)]
    assert [r[2] for r in bucket_stage_ranges(12)] == list(range(11, 0, -1)) + [-1]
