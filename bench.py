#!/usr/bin/env python
"""Benchmark of the RoViT-KAN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode infer|train|kan|sweep] [--batch B] [--impl ours|reference]

Default = BASELINE.json configs[1]: RoViT-KAN inference, batch 1024 per GPU, bf16 tensor-core trunk, all four
heads (KAN severity enabled), random-init weights, synthetic 224x224 images.  A "step" is one forward pass
of the whole model over one batch.  With --gpus N > 1 (launched by torch.distributed.run) every rank runs its
own batch: the path is data parallel with no data-path collective in inference ("scaling": "weak"); in
--mode train ranks all-reduce the flat gradient once per step over NCCL.

Timing: W untimed steps, then K steps bracketed by barrier + cuda synchronize, CUDA events on the launch
stream, max over ranks.  The 1024-image input (616 MB fp32) and every inter-kernel tensor set exceed the
126 MB L2, so no L2 flush is needed between iterations ("l2": "inputs larger than L2").

One JSON line is printed by rank 0; see the task contract for the keys.  `roofline` describes the dominant
kernel (inference: mlp_fused_kernel, one launch per block, ~50 % of the step; training: the tcgen05 GEMM family); its
per-launch device time is measured in-situ with CUDA events on the launch stream by librovitkan (rvk_gemm_timing_*)
in K extra steps after the timed region; `traffic` (DRAM bytes per launch) comes from the committed ncu --set full
capture of the same kernel (profiles/roofline_traffic.json).
`cpu_baseline` / `--impl reference` time the reference's CPU algorithm (oracle port incl. the reference's
per-(input,output) Python loop in the KAN, models/kan.py:85-89) on the host cores, batch 32 per step.
"""

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_FLOP_PER_IMG = 2.507e9          # 1253.7 M MAC per image, forward
TRAIN_FLOP_PER_IMG = 7.52e9


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tflops_sustained': d['bf16_tflops_sustained'], 'tflops_burst': d['bf16_tflops'],
                'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops_sustained': 1400.0, 'tflops_burst': 1590.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [x.strip() for x in l.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(steps, warmup, batch=32, quiet=False):
    """Reference algorithm on the host cores: oracle port with the reference's KAN double loop."""
    import torch
    from oracle import model as omodel
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = omodel.random_state_dict(0)
    x = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        for _ in range(warmup):
            omodel.forward(sd, x, stage=4, kan_loop=True)
        t0 = time.perf_counter()
        for _ in range(steps):
            omodel.forward(sd, x, stage=4, kan_loop=True)
        dt = time.perf_counter() - t0
    return {'value': batch * steps / dt, 'unit': 'images/sec', 'cores': cores, 'kind': 'port',
            'sample': f'{steps} eval forwards of batch {batch} (fp32, stage 4, KAN via the reference\'s per-(input,output) '
                      f'loop) after {warmup} warm-up; {dt:.1f} s of CPU work', 'ms_per_step': dt / steps * 1e3}


def cpu_reference_train(steps, warmup, batch=8):
    """Reference train step (forward + joint loss + backward through torch autograd, no optimizer) on the host cores."""
    import torch
    from oracle import losses as olosses
    from oracle import model as omodel
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'knots' not in k else v)
          for k, v in omodel.random_state_dict(0).items()}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, 3, 224, 224, generator=g)
    y = torch.randint(0, 4, (batch,), generator=g)

    def step():
        for v in sd.values():
            if v.requires_grad:
                v.grad = None
        olosses.joint(omodel.forward(sd, x, stage=4, kan_loop=True), y, y, 4)['total_loss'].backward()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {'value': batch * steps / dt, 'unit': 'images/sec', 'cores': cores, 'kind': 'port',
            'sample': f'{steps} stage-4 train steps (forward + joint loss + backward, fp32, KAN via the reference\'s per-(input,output) '
                      f'loop) of batch {batch} after {warmup} warm-up; {dt:.1f} s of CPU work', 'ms_per_step': dt / steps * 1e3}


def workload_config(train, batch, world):
    """The `config` object both arms print (the reference arm times a bounded sample of the same workload)."""
    return {'workload': ('RoViT-KAN curriculum stage-4 training step (all losses, AdamW), batch %d/GPU' % batch) if train
            else 'RoViT-KAN inference, batch %d per GPU, bf16 tensor-core trunk, all four heads (KAN severity enabled)' % batch,
            'batch_per_gpu': batch, 'global_batch': batch * world, 'image': '3x224x224', 'chunk_images': min(batch, 2048),
            'parallelism': f'dp{world}', 'l2': 'inputs larger than L2 (616 MB per batch), no flush needed',
            'weights': 'random init (timm init laws), seed 0'}


def metric_name(train):
    return 'images/sec (224^2, device-timed) RoViT-KAN ' + ('train step' if train else 'inference forward')


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path (oracle port: the reference is a Python project whose trunk lives in
    the absent `timm`), all host threads, on a bounded sample of OUR arm's workload; same metric / unit / config keys."""
    if rank != 0:
        return
    train = args.mode == 'train'
    if train:
        steps, warmup = max(1, min(args.steps, 3)), 1
        r = cpu_reference_train(steps, warmup)
    else:
        steps, warmup = max(1, args.steps), max(1, min(args.warmup, 2))
        r = cpu_reference(steps, warmup)
    batch = args.batch or (256 if train else 1024)
    cfg = workload_config(train, batch, max(1, args.gpus))
    cfg['reference_sample'] = r['sample']
    line = {'impl': 'reference', 'metric': metric_name(train), 'value': r['value'],
            'unit': 'images/sec', 'n_gpus': args.gpus, 'steps': steps, 'warmup': warmup, 'ms_per_step': r['ms_per_step'],
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': cfg,
            'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
            'e2e': {'value': r['value'], 'unit': 'images/sec', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from rovitkan_b200 import _lib
    from rovitkan_b200.models import RoViTKAN
    from rovitkan_b200.training.losses import JointLoss

    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = _lib.load()
    train = args.mode == 'train'
    batch = args.batch or (256 if train else 1024)
    K, W = args.steps, max(3, args.warmup)

    torch.manual_seed(0)
    model = RoViTKAN(pretrained=False).to(dev)
    g = torch.Generator().manual_seed(1000 + rank)
    host_images = torch.randn(batch, 3, 224, 224, generator=g).pin_memory()
    labels = torch.randint(0, 4, (batch,), generator=g)
    images = host_images.to(dev)
    yd = labels.to(dev)

    if train:
        model.train()
        loss_fn = JointLoss(focal_alpha=torch.ones(4))
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4, fused=True)
        params = [p for p in model.parameters()]

        def step(x):
            out = model(x)
            loss = loss_fn(out, yd, yd, 4)['total_loss']
            opt.zero_grad(set_to_none=True)
            loss.backward()
            if world > 1:
                from rovitkan_b200.dist import all_reduce_gradients
                all_reduce_gradients(params, world)
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            return loss
    else:
        model.eval()

        def step(x):
            with torch.no_grad():
                return model(x)['kan_severity']

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    for _ in range(W):
        step(images)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = lib.rvk_launch_count()
    ms_total = timed(lambda: step(images), K)
    launches = lib.rvk_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    value = batch * world * K / (ms_total / 1e3)

    # end-to-end through the public API: every step copies ITS batch from pinned host memory to the device and
    # reads its result back to the host, all inside the timed region.  The copy of step i+1 runs on a second
    # stream while step i computes (double-buffered), as a serving loop would do it.
    copy_stream = torch.cuda.Stream(device=dev)

    def make_e2e(host_batch):
        bufs = [torch.empty(host_batch.shape, dtype=host_batch.dtype, device=dev) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def e2e_run(n):
            main = torch.cuda.current_stream()
            for e in consumed:
                e.record(main)

            def prefetch(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[i % 2])
                    bufs[i % 2].copy_(host_batch, non_blocking=True)
                    ready[i % 2].record(copy_stream)
            prefetch(0)
            last = None
            for i in range(n):
                if i + 1 < n:
                    prefetch(i + 1)
                main.wait_event(ready[i % 2])
                r = step(bufs[i % 2])
                consumed[i % 2].record(main)
                last = r.float().cpu()           # device -> host read of this step's result (synchronises)
            return last
        return e2e_run

    e2e_run = make_e2e(host_images)

    def timed_run(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(n)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    e2e_run(2)
    ms_e2e = timed_run(e2e_run, K)
    e2e_value = batch * world * K / (ms_e2e / 1e3)
    # serving variant: the host batch is already bf16 (the trunk rounds pixels to bf16 first, so the results are
    # bit-identical); halves the PCIe bytes that bound the fp32 end-to-end number
    e2e_bf16 = None
    if not train:
        host_bf16 = host_images.to(torch.bfloat16).pin_memory()
        run_bf16 = make_e2e(host_bf16)
        run_bf16(2)
        ms_b = timed_run(run_bf16, K)
        e2e_bf16 = {'value': batch * world * K / (ms_b / 1e3), 'unit': 'images/sec', 'h2d_bytes_per_step': host_bf16.numel() * 2,
                    'd2h_bytes_per_step': batch * 4, 'ms_per_step': ms_b / K}
    d2h = 4 if train else batch * 4

    # in-situ device time of the dominant kernel family (tcgen05 GEMMs), K more steps
    lib.rvk_gemm_timing_enable(1)
    torch.cuda.synchronize()
    for _ in range(K):
        step(images)
    torch.cuda.synchronize()
    t_ms, t_fl = ctypes.c_double(0), ctypes.c_double(0)
    n_gemm = lib.rvk_gemm_timing_collect(ctypes.byref(t_ms), ctypes.byref(t_fl))
    lib.rvk_gemm_timing_enable(0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    gemm_tflops = (t_fl.value / (t_ms.value * 1e-3) / 1e12) if t_ms.value > 0 else 0.0
    flop_img = TRAIN_FLOP_PER_IMG if train else FWD_FLOP_PER_IMG
    # dominant kernel: the fused (proj +) MLP kernel in inference (~50 % of the step), the dgrad/wgrad GEMM family in training
    kinds = {}
    for kind, name in ((0, 'gemm_nt_kernel'), (1, 'gemm_tn_kernel'), (2, 'mlp_fused_kernel')):
        k_ms, k_fl = ctypes.c_double(0), ctypes.c_double(0)
        n = lib.rvk_gemm_timing_kind(kind, ctypes.byref(k_ms), ctypes.byref(k_fl))
        if n > 0 and k_ms.value > 0:
            kinds[name] = {'launches': int(n), 'us_per_launch': k_ms.value * 1e3 / n, 'gflop_per_launch': k_fl.value / n / 1e9,
                           'tflops': k_fl.value / (k_ms.value * 1e-3) / 1e12, 'ms_per_step': k_ms.value / K}
    if not train and 'mlp_fused_kernel' in kinds:
        dom = kinds['mlp_fused_kernel']
        dom_name = ('mlp_fused_kernel<2> (attention out-projection + LayerNorm2 + fc1 + GELU + fc2 + residual + LayerNorm1, '
                    'one launch per block; algorithmic FLOPs 2*M*192*(192 + 2*768) per launch, M = batch*197)')
        dom_tflops = dom['tflops']
    else:
        dom, dom_name, dom_tflops = None, 'gemm_nt_kernel / gemm_tn_kernel / mlp_fused_kernel (tcgen05, all launches)', gemm_tflops
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')) as f:
            tj = json.load(f)
        ent = tj.get('train' if train else 'infer')
        if ent and ent.get('batch_per_gpu') == batch:
            traffic, traffic_src = ent['dram_bytes_per_launch'], ent['source']
    except (OSError, ValueError, KeyError):
        pass
    line = {
        'metric': metric_name(train),
        'value': value, 'unit': 'images/sec', 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms_total / K,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': workload_config(train, batch, world),
        'e2e': {'value': e2e_value, 'unit': 'images/sec', 'h2d_bytes_per_step': host_images.numel() * 4,
                'd2h_bytes_per_step': d2h, 'ms_per_step': ms_e2e / K},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'e2e_bf16_input': e2e_bf16,
        'roofline': {'bound': 'tensor', 'achieved': dom_tflops, 'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s',
                     'frac': dom_tflops / pk['tflops_sustained'], 'traffic': traffic, 'traffic_source': traffic_src,
                     'kernel': dom_name,
                     'us_per_launch': dom['us_per_launch'] if dom else None,
                     'gflop_per_launch': dom['gflop_per_launch'] if dom else None,
                     'share_of_step': (dom['ms_per_step'] / (ms_total / K)) if dom else None,
                     'launches_timed': int(dom['launches'] if dom else n_gemm),
                     'peak_source': pk['source'] + ' sustained bf16',
                     'kernels': kinds,
                     'all_tcgen05_gemms': {'tflops': gemm_tflops, 'frac': gemm_tflops / pk['tflops_sustained'],
                                           'ms_per_step': t_ms.value / K, 'launches_timed': int(n_gemm)},
                     'whole_step_tflops': value / world * flop_img / 1e12,
                     'whole_step_frac': value / world * flop_img / 1e12 / pk['tflops_sustained']},
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_train(steps=3, warmup=1) if train else cpu_reference(steps=args.cpu_steps, warmup=1)
        line['cpu_baseline'] = {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------ KAN microbenchmark
def run_kan(args, rank, local_rank, world):
    """BASELINE.json configs[2]: KANSeverityModule([192,64,1]) (and the production [192,64,16,1]), grid=5 k=3, batch
    65536, forward + backward (dx, dW, dWl, db of every layer).  Algorithmic bytes / FLOPs per fwd+bwd from
    SURVEY.md section 8(d): 152 MB, 38.7 GFLOP for the [192,64,1] stack."""
    import torch
    from rovitkan_b200 import _lib
    from rovitkan_b200.models.kan import KANSeverityModule
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    lib = _lib.load()
    batch = args.batch or 65536
    K, W = args.steps, max(3, args.warmup)
    out = {}
    for name, dims in (('192-64-1', [192, 64, 1]), ('192-64-16-1', [192, 64, 16, 1])):
        torch.manual_seed(0)
        m = KANSeverityModule(dims).to(dev)
        x = torch.randn(batch, 192, generator=torch.Generator().manual_seed(0)).to(dev).requires_grad_(True)
        gy = torch.randn(batch, 1, generator=torch.Generator().manual_seed(1)).to(dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        def fwd():
            with torch.no_grad():
                return m(x)

        def fwdbwd():
            y = m(x)
            y.backward(gy)
            m.zero_grad(set_to_none=True)
            x.grad = None

        res = {}
        for tag, fn in (('fwd', fwd), ('fwd_bwd', fwdbwd)):
            for _ in range(W):
                fn()
            torch.cuda.synchronize()
            tot = 0.0
            l0 = lib.rvk_launch_count()
            for _ in range(K):
                flush.zero_()                      # L2 flush between timed iterations (x is 50 MB < 126 MB L2)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            res[tag] = {'ms': tot / K, 'launches': (lib.rvk_launch_count() - l0) // K}
        out[name] = res
    if rank != 0:
        return
    pk = peaks()
    r = out['192-64-1']
    ms = r['fwd_bwd']['ms']
    alg_bytes, alg_flops = 152.0e6 * batch / 65536, 38.7e9 * batch / 65536
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    line = {
        'metric': 'samples/sec KANSeverityModule([192,64,1]) forward+backward', 'value': batch / (ms * 1e-3), 'unit': 'samples/sec',
        'n_gpus': 1, 'steps': K, 'warmup': W, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'KANLayer microbench: 192->64->1 spline head, grid=5 k=3, batch %d fwd+bwd' % batch,
                   'l2': 'L2 flushed (256 MB memset) between timed iterations'},
        'gpu_launches': int(r['fwd_bwd']['launches']) * K,
        'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': gbs / pk['hbm_gbs'],
                     'traffic': None,
                     'kernel': 'kan_fwd_tc_kernel + kan_bwd_x_tc_kernel + kan_bwd_w_tc_kernel (tcgen05, operands generated on the fly, '
                               'hi+lo bf16 split = 3 MMAs per product) for 192->64; fp32 CUDA-core kernels for 64->1',
                     'note': 'algorithmic 152 MB / 38.7 GFLOP per fwd+bwd (%.1f TFLOP/s dense-equivalent achieved): the three kernels '
                             'are bound by the CUDA-core generation of the expanded activations (tanh, interval search, four cubics, '
                             'hi/lo split: ~100 instructions per (sample, input), done once per kernel), not by HBM or the tensor pipe'
                             % (alg_flops / (ms * 1e-3) / 1e12)},
        'detail': out,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ throughput sweep
def run_sweep(args, rank, local_rank, world):
    """BASELINE.json configs[4]: forward throughput of DeiT-Tiny backbone + heads + KAN at batch 64..8192 per GPU, each
    rank on its own shard (no collective), against the tensor roofline (2.507 GFLOP per image)."""
    import torch
    import torch.distributed as dist
    from rovitkan_b200.models import RoViTKAN
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(0)
    model = RoViTKAN(pretrained=False).to(dev).eval()
    K, W = args.steps, max(3, args.warmup)
    pk = peaks()
    rows = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > L2: small batches would otherwise stay L2-resident
    for batch in (64, 128, 148, 256, 296, 512, 1024, 2048, 4096, 8192):
        g = torch.Generator(device=dev).manual_seed(1000 + rank)
        images = torch.randn(batch, 3, 224, 224, generator=g, device=dev)
        with torch.no_grad():
            for _ in range(W):
                model(images)
            tot = 0.0
            for _ in range(K):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                e0.record()
                model(images)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
        ms = torch.tensor([tot / K], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ips = batch * world / (float(ms) / 1e3)
        rows.append({'batch_per_gpu': batch, 'ms_per_step': float(ms), 'images_per_sec': ips,
                     'tensor_roofline_frac': ips / world * FWD_FLOP_PER_IMG / 1e12 / pk['tflops_sustained']})
        del images
    if rank == 0:
        print(json.dumps({'metric': 'images/sec (224^2, device-timed) RoViT-KAN inference forward, batch sweep', 'unit': 'images/sec',
                          'n_gpus': world, 'steps': K, 'warmup': W, 'higher_is_better': True, 'scaling': 'weak', 'dtype': 'bf16',
                          'data': 'synthetic', 'config': {'workload': 'encoder-throughput sweep, batch 64-8192 per GPU', 'parallelism': f'dp{world}',
                                                          'l2': 'L2 flushed (256 MB memset) before every timed step'},
                          'peak_tflops': pk['tflops_sustained'], 'sweep': rows}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='infer', choices=['infer', 'train', 'kan', 'sweep'])
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-steps', type=int, default=40, help='batch-32 reference forwards timed for cpu_baseline (~10-20 s)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    elif args.mode == 'kan':
        if rank == 0:
            run_kan(args, rank, local_rank, world)
    elif args.mode == 'sweep':
        run_sweep(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
