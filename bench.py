#!/usr/bin/env python
"""Benchmark of the RoViT-KAN hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode all|infer|train|kan|sweep] [--batch B] [--impl ours|reference]

Default (`--mode all`) prints ONE JSON line whose headline (`metric`, `value`, `e2e`, `roofline`, ...) is BASELINE.json
configs[1] -- RoViT-KAN inference, batch 1024 per GPU, bf16 tensor-core trunk, all four heads (KAN severity enabled),
random-init weights, synthetic 224x224 images -- and which carries the other BASELINE configs as sub-objects measured in the
same process at the same N:
    "train"  configs[3]  stage-4 training step (all losses, AdamW with the reference's two LR groups), batch 256 per GPU,
                         gradient all-reduce over NCCL when N > 1; phases, per-kernel rooflines, CutMix and frozen-backbone variants
    "kan"    configs[2]  KANSeverityModule([192,64,1]) (and the production [192,64,16,1]), batch 65536, forward + backward
    "sweep"  configs[4]  forward throughput at batch 64..8192 per GPU against the tensor roofline
    "gpu_eager_baseline" the oracle restatement of the reference (PyTorch eager: cuBLAS / SDPA kernels) on the same B200
`--mode infer|train|kan|sweep` runs one of them alone as its own headline line.

A "step" is one pass of the path over one batch.  With --gpus N > 1 (launched by torch.distributed.run) every rank runs
its own batch: pure data parallel, "scaling": "weak"; the only collective is the gradient all-reduce of the train step.

Timing: W untimed steps, then K steps bracketed by barrier + cuda synchronize, CUDA events on the launch stream, max over
ranks.  The 1024-image input (616 MB fp32) and every inter-kernel tensor set exceed the 126 MB L2, so the inference and
train steps need no L2 flush ("l2" in `config`); the KAN microbenchmark and the sweep flush L2 between timed iterations.

Per-kernel numbers (`roofline.kernels`, `train.kernels`, `kan.kernels`): device time of every launch of a kernel family,
CUDA events on the launch stream inside librovitkan (rvk_timing_*), measured in K extra steps after the timed region;
`traffic` (DRAM bytes per launch) comes from the committed ncu --set full captures (profiles/roofline_traffic.json).
`cpu_baseline` / `--impl reference`: the reference's CPU algorithm (oracle port incl. the reference's per-(input,output)
Python loop in the KAN, models/kan.py:85-89) on the host cores.
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
KAN_BYTES_PER_SAMPLE = 152.0e6 / 65536      # [192,64,1] fwd+bwd algorithmic bytes
KAN_FLOP_PER_SAMPLE = 38.7e9 / 65536
SWEEP_BATCHES = (64, 128, 148, 256, 296, 512, 1024, 2048, 4096, 8192)
# which roofline bounds each kernel family (rvk_timing_kind_name): tensor pipe for the tcgen05 kernels, HBM otherwise
TENSOR_BOUND = {'gemm_nt_kernel', 'gemm_tn_kernel', 'mlp_fused_kernel', 'attn_fwd_tc_kernel', 'attn_bwd_tc_kernel'}


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


# ------------------------------------------------------------------------------------------ CPU reference legs
def cpu_reference(steps, warmup, batch=32):
    """Reference algorithm on the host cores: oracle port with the reference's KAN double loop; eval forwards."""
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
    return {'value': batch * steps / dt, 'unit': 'images/sec', 'cores': cores, 'kind': 'port', 'batch': batch,
            'sample': f'{steps} eval forwards of batch {batch} (fp32, stage 4, KAN via the reference\'s per-(input,output) '
                      f'loop) after {warmup} warm-up; {dt:.1f} s of CPU work', 'ms_per_step': dt / steps * 1e3}


def cpu_reference_train(steps, warmup, batch=32):
    """Reference train step (forward + joint loss + backward through torch autograd, no optimizer) on the host cores,
    batch 32 = the reference's own training batch (configs/config.py:33)."""
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
    return {'value': batch * steps / dt, 'unit': 'images/sec', 'cores': cores, 'kind': 'port', 'batch': batch,
            'sample': f'{steps} stage-4 train steps (forward + joint loss + backward, fp32, KAN via the reference\'s per-(input,output) '
                      f'loop) of batch {batch} after {warmup} warm-up; {dt:.1f} s of CPU work', 'ms_per_step': dt / steps * 1e3}


def cpu_reference_kan(batches=(32,), dims=(192, 64, 1)):
    """Reference KANSeverityModule forward + backward (per-(input,output) loop port, autograd) on the host cores."""
    import torch
    from oracle import kan as okan
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = []
    for b in batches:
        g = torch.Generator().manual_seed(0)
        layers = [tuple(p.requires_grad_(True) for p in l) for l in okan.init_layers(list(dims), generator=g)]
        x = torch.randn(b, dims[0], generator=g).requires_grad_(True)
        gy = torch.randn(b, 1, generator=g)
        t0 = time.perf_counter()
        y = okan.severity_forward(x, layers, okan.make_knots(), loop=True)
        t1 = time.perf_counter()
        y.backward(gy)
        t2 = time.perf_counter()
        out.append({'batch': b, 'fwd_ms': (t1 - t0) * 1e3, 'fwd_bwd_ms': (t2 - t0) * 1e3, 'samples_per_sec': b / (t2 - t0)})
    return {'kind': 'port', 'cores': cores, 'unit': 'samples/sec', 'value': out[0]['samples_per_sec'], 'runs': out,
            'sample': f'KANSeverityModule({list(dims)}) forward + backward through the reference\'s per-(input,output) loop, one '
                      f'run each at batch {list(batches)} (loop cost is per batch, not per sample)'}


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
    the absent `timm`), all host threads, on a bounded sample of OUR arm's workload; same metric / unit / config / steps /
    warm-up as our arm.  One step = one batch-32 forward (1/32 of the batch-1024 step; BASELINE configs[0]); because the
    reference's KAN loop costs per BATCH, one full batch-1024 forward is timed too and reported as `same_batch`."""
    if rank != 0:
        return
    train = args.mode == 'train'
    K, W = args.steps, max(3, args.warmup)
    batch = args.batch or (256 if train else 1024)
    line = {'impl': 'reference', 'metric': metric_name(train), 'unit': 'images/sec', 'n_gpus': args.gpus,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32', 'data': 'synthetic',
            'config': workload_config(train, batch, max(1, args.gpus)), 'gpu_launches': 0}
    if train:
        steps, warmup = max(1, min(K, 2)), 1
        r = cpu_reference_train(steps, warmup)
    else:
        steps, warmup = K, W
        r = cpu_reference(steps, warmup)
        full = cpu_reference(1, 0, batch=batch)
        line['same_batch'] = {'batch': batch, 'value': full['value'], 'unit': 'images/sec', 'ms_per_step': full['ms_per_step'],
                              'sample': full['sample']}
        if args.mode == 'all' and args.gpus == 1 and not args.no_subs:
            t = cpu_reference_train(1, 1)
            line['train'] = {'value': t['value'], 'unit': 'images/sec', 'ms_per_step': t['ms_per_step'], 'sample': t['sample']}
            line['kan'] = cpu_reference_kan()
    line.update({'value': r['value'], 'steps': steps, 'warmup': warmup, 'ms_per_step': r['ms_per_step'],
                 'reference_sample': r['sample'],
                 'cpu_baseline': {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
                 'e2e': {'value': r['value'], 'unit': 'images/sec', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}})
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ our arm: shared plumbing
class Ctx:
    def __init__(self, rank, local_rank, world):
        import torch
        import torch.distributed as dist
        from rovitkan_b200 import _lib
        assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
        self.torch, self.dist = torch, dist
        self.rank, self.local_rank, self.world = rank, local_rank, world
        torch.cuda.set_device(local_rank)
        self.dev = torch.device('cuda', local_rank)
        if world > 1 and not dist.is_initialized():
            dist.init_process_group('nccl', device_id=self.dev)
        self.lib = _lib.load()
        self.pk = peaks()
        self._flush = None
        self.numa = self._gpu_local_cpus()

    def _gpu_local_cpus(self):
        """CPUs of the NUMA node this rank's GPU hangs off (NVML), so that the pinned staging buffers of the e2e loops are
        allocated on that node: with 8 ranks on a two-socket host every buffer on one node would funnel half of the
        host->device traffic through the socket interconnect."""
        info = {'bound': False}
        try:
            import pynvml
            pynvml.nvmlInit()
            prop = self.torch.cuda.get_device_properties(self.local_rank)
            try:
                bus = '%08x:%02x:%02x.0' % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
                h = pynvml.nvmlDeviceGetHandleByPciBusId(bus)
            except Exception:
                vis = os.environ.get('CUDA_VISIBLE_DEVICES', '')
                ids = [v for v in vis.split(',') if v.strip().isdigit()]
                h = pynvml.nvmlDeviceGetHandleByIndex(int(ids[self.local_rank]) if ids else self.local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
            info['gpu_cpus'] = len(cpus)
            try:
                info['numa_node'] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
            except Exception:
                pass
            self._local_cpus = cpus & os.sched_getaffinity(0)
            info['bound'] = bool(self._local_cpus) and len(self._local_cpus) < len(os.sched_getaffinity(0))
        except Exception as e:                         # NVML absent / container without topology: allocate as before
            self._local_cpus = set()
            info['error'] = type(e).__name__
        return info

    def pinned(self, t):
        """Pinned host copy of `t`, allocated (cudaHostAlloc places the pages) and filled while this thread runs on the
        GPU's own NUMA node; the affinity is restored afterwards (the CPU baseline legs use every core)."""
        old = os.sched_getaffinity(0)
        try:
            if self._local_cpus and os.environ.get('RVK_BENCH_NUMA', '1') != '0':
                os.sched_setaffinity(0, self._local_cpus)
            out = self.torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            out.copy_(t)
            return out
        finally:
            os.sched_setaffinity(0, old)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world == 1:
            return float(ms)
        t = self.torch.tensor([float(ms)], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def timed(self, fn, n):
        """n calls of fn() between barrier + synchronize, CUDA events on the launch stream, max over ranks -> ms total."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def timed_run(self, fn, n):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(n)
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def flush_l2(self):
        if self._flush is None:
            self._flush = self.torch.empty(256 << 20, dtype=self.torch.uint8, device=self.dev)
        self._flush.zero_()

    def kernel_table(self, fn, n, step_ms):
        """Device time of every kernel family of librovitkan over n more calls of fn(): CUDA events around each launch
        on its own stream (this serialises programmatic dependent launches, so the sum can exceed the step time)."""
        lib, pk = self.lib, self.pk
        lib.rvk_timing_enable(1)
        lib.rvk_set_side_stream(0)          # every kernel timed on its own: no concurrent weight-gradient GEMMs next to it
        self.torch.cuda.synchronize()
        for _ in range(n):
            fn()
        self.torch.cuda.synchronize()
        lib.rvk_timing_collect()
        lib.rvk_timing_enable(0)
        lib.rvk_set_side_stream(-1)
        table, kind = {}, 0
        while True:
            name = lib.rvk_timing_kind_name(kind)
            if name is None:
                break
            ms, fl, by = ctypes.c_double(0), ctypes.c_double(0), ctypes.c_double(0)
            cnt = lib.rvk_timing_kind(kind, ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(by))
            kind += 1
            if cnt <= 0 or ms.value <= 0:
                continue
            name = name.decode()
            e = {'launches': int(cnt), 'launches_per_step': cnt / n, 'us_per_launch': ms.value * 1e3 / cnt,
                 'ms_per_step': ms.value / n, 'share_of_step': ms.value / n / step_ms}
            if fl.value > 0:
                e['gflop_per_launch'] = fl.value / cnt / 1e9
                e['tflops'] = fl.value / (ms.value * 1e-3) / 1e12
                e['frac_tensor'] = e['tflops'] / pk['tflops_sustained']
            if by.value > 0:
                e['mb_per_launch'] = by.value / cnt / 1e6
                e['gbs'] = by.value / (ms.value * 1e-3) / 1e9
                e['frac_hbm'] = e['gbs'] / pk['hbm_gbs']
            e['bound'] = 'tensor' if name in TENSOR_BOUND else 'hbm'
            e['frac'] = e.get('frac_tensor' if e['bound'] == 'tensor' else 'frac_hbm')
            table[name] = e
        return table

    def make_e2e(self, step, host_batch):
        """End-to-end loop through the public nn.Module call: every step copies ITS batch from pinned host memory to the
        device and reads its result back to the host, all inside the timed region.  The copy of step i+1 runs on a second
        stream while step i computes (double-buffered), as a serving / training input loop would do it."""
        torch, dev = self.torch, self.dev
        copy_stream = torch.cuda.Stream(device=dev)
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

    def e2e(self, step, host_batch, K, batch, d2h_bytes):
        run = self.make_e2e(step, host_batch)
        run(2)
        ms = self.timed_run(run, K)
        return {'value': batch * self.world * K / (ms / 1e3), 'unit': 'images/sec',
                'h2d_bytes_per_step': host_batch.numel() * host_batch.element_size(), 'd2h_bytes_per_step': d2h_bytes,
                'ms_per_step': ms / K, 'input_dtype': str(host_batch.dtype).replace('torch.', ''),
                'pinned_on_gpu_numa_node': self.numa}


def traffic_entry(key, batch):
    try:
        with open(os.path.join(ROOT, 'profiles', 'roofline_traffic.json')) as f:
            ent = json.load(f).get(key)
        if ent and ent.get('batch_per_gpu') == batch:
            return ent['dram_bytes_per_launch'], ent['source']
    except (OSError, ValueError, KeyError):
        pass
    return None, None


# ------------------------------------------------------------------------------------------ configs[1]: inference
def bench_infer(ctx, K, W, batch, with_e2e=True):
    torch = ctx.torch
    from rovitkan_b200.models import RoViTKAN
    torch.manual_seed(0)
    model = RoViTKAN(pretrained=False).to(ctx.dev).eval()
    g = torch.Generator().manual_seed(1000 + ctx.rank)
    host_images = ctx.pinned(torch.randn(batch, 3, 224, 224, generator=g))
    images = host_images.to(ctx.dev)

    def step(x):
        with torch.no_grad():
            return model(x)['kan_severity']
    for _ in range(W):
        step(images)
    sampler = ClockSampler(ctx.local_rank)
    if ctx.rank == 0:
        sampler.start()
    l0 = ctx.lib.rvk_launch_count()
    ms_total = ctx.timed(lambda: step(images), K)
    launches = ctx.lib.rvk_launch_count() - l0
    clocks = sampler.stop() if ctx.rank == 0 else None
    value = batch * ctx.world * K / (ms_total / 1e3)
    step_ms = ms_total / K
    res = {'value': value, 'ms_per_step': step_ms, 'gpu_launches': int(launches), 'clocks': clocks, 'batch': batch}
    if with_e2e:
        res['e2e'] = ctx.e2e(step, host_images, K, batch, batch * 4)
        # serving variants: the host batch already bf16 (bit-identical results: the trunk rounds pixels to bf16 first) or
        # uint8 pixels as an image decoder produces them (ToTensor + Normalize folded into the patch gather)
        res['e2e_bf16_input'] = ctx.e2e(step, ctx.pinned(host_images.to(torch.bfloat16)), K, batch, batch * 4)
        host_u8 = ctx.pinned(torch.randint(0, 256, (batch, 3, 224, 224), dtype=torch.uint8, generator=g))
        res['e2e_uint8_input'] = ctx.e2e(step, host_u8, K, batch, batch * 4)
    kernels = ctx.kernel_table(lambda: step(images), K, step_ms)
    res['kernels'] = kernels
    pk = ctx.pk
    dom = kernels.get('mlp_fused_kernel')
    traffic, traffic_src = traffic_entry('infer', batch)
    tc = [e for n, e in kernels.items() if n in TENSOR_BOUND]
    tc_ms = sum(e['ms_per_step'] for e in tc)
    tc_fl = sum(e['gflop_per_launch'] * e['launches_per_step'] for e in tc) * 1e9
    res['roofline'] = {
        'bound': 'tensor', 'achieved': dom['tflops'] if dom else None, 'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s',
        'frac': dom['frac_tensor'] if dom else None, 'traffic': traffic, 'traffic_source': traffic_src,
        'kernel': ('mlp_fused2_kernel (CTA pairs, two row tiles in flight; RVK_MLP_CTA_GROUP=2: mlp_fused_kernel<2>): attention '
                   'out-projection + LayerNorm2 + fc1 + GELU + fc2 + residual + LayerNorm1, '
                   'one launch per block; algorithmic FLOPs 2*M*192*(192 + 2*768) per launch, M = batch*197'),
        'us_per_launch': dom['us_per_launch'] if dom else None, 'gflop_per_launch': dom['gflop_per_launch'] if dom else None,
        'share_of_step': dom['share_of_step'] if dom else None, 'launches_timed': dom['launches'] if dom else 0,
        'peak_source': pk['source'] + ' sustained bf16', 'kernels': kernels,
        'all_tcgen05_kernels': {'tflops': tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else None, 'ms_per_step': tc_ms,
                                'frac': tc_fl / (tc_ms * 1e-3) / 1e12 / pk['tflops_sustained'] if tc_ms > 0 else None},
        'whole_step_tflops': value / ctx.world * FWD_FLOP_PER_IMG / 1e12,
        'whole_step_frac': value / ctx.world * FWD_FLOP_PER_IMG / 1e12 / pk['tflops_sustained']}
    del model, images, host_images
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------ configs[3]: train step
def build_optimizer(model, fused_tail=True):
    """The reference's optimizer (training/optimizer.py:7-32): AdamW, backbone parameters at lr/10, weight decay 1e-4."""
    import torch
    backbone = [p for n, p in model.named_parameters() if 'backbone' in n and p.requires_grad]
    heads = [p for n, p in model.named_parameters() if 'backbone' not in n and p.requires_grad]
    groups = [{'params': backbone, 'lr': 1e-5}, {'params': heads, 'lr': 1e-4}]
    groups = [g for g in groups if g['params']]
    if fused_tail:
        try:
            from rovitkan_b200.training.optim import FusedAdamW
            return FusedAdamW(groups, weight_decay=1e-4, max_grad_norm=1.0), True
        except ImportError:
            pass
    return torch.optim.AdamW(groups, weight_decay=1e-4, fused=True), False


def bench_train(ctx, K, W, batch, with_variants=True):
    torch, dist = ctx.torch, ctx.dist
    from rovitkan_b200.dist import all_reduce_gradients
    from rovitkan_b200.models import RoViTKAN
    from rovitkan_b200.training.losses import JointLoss
    torch.manual_seed(0)
    model = RoViTKAN(pretrained=False).to(ctx.dev).train()
    g = torch.Generator().manual_seed(2000 + ctx.rank)
    host_images = ctx.pinned(torch.randn(batch, 3, 224, 224, generator=g))
    yd = torch.randint(0, 4, (batch,), generator=g).to(ctx.dev)
    yb = yd.flip(0)
    images = host_images.to(ctx.dev)
    loss_fn = JointLoss(focal_alpha=torch.ones(4))
    opt, fused_tail = build_optimizer(model, not os.environ.get('RVK_TORCH_OPTIMIZER'))
    params = [p for p in model.parameters()]
    ev = {}
    from rovitkan_b200 import dist as rdist
    # default: ONE in-place all-reduce of the trunk's flat gradient + one of the heads' after the backward pass.  Bucketed
    # all-reduce launched from inside the backward (RVK_DP_OVERLAP=1) hides ~0.04 ms of a 0.12 ms transfer but its NCCL CTAs
    # run next to the backward kernels: measured 322.6 k vs 326.9 k img/s at 8 GPUs, 74.0 k vs 74.4 k at 2 -- both are
    # measured below (`allreduce.overlap_ab`) at every N > 1
    want_overlap = os.environ.get('RVK_DP_OVERLAP') == '1'
    overlap = ctx.world > 1 and want_overlap and rdist.enable_overlap(int(os.environ.get('RVK_DP_BUCKETS', '3')))
    average = True
    if ctx.world > 1 and fused_tail:
        opt.grad_mult, average = 1.0 / ctx.world, False      # the 1/world scale rides in the optimizer kernel

    def mark(name):
        if ev.get('on'):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.setdefault('marks', []).append((name, e))

    def step(x, cutmix=False):
        mark('start')
        out = model(x)
        if cutmix:        # trainer.py:104-111: both label sets through the loss, blended by lam
            la, lb = loss_fn(out, yd, yd, 4), loss_fn(out, yb, yd, 4)
            loss = 0.7 * la['total_loss'] + 0.3 * lb['total_loss']
        else:
            loss = loss_fn(out, yd, yd, 4)['total_loss']
        opt.zero_grad(set_to_none=True)
        loss.backward()
        mark('backward_done')
        if ctx.world > 1:
            ev['collectives'] = all_reduce_gradients(params, ctx.world, average=average)
        mark('allreduce_done')
        if fused_tail:
            opt.step()                        # unscale + global-norm clip + AdamW + weight shadows in one pass
        else:
            torch.nn.utils.clip_grad_norm_([p for p in params if p.grad is not None], 1.0)
            opt.step()
        mark('optimizer_done')
        return loss

    for _ in range(W):
        step(images)
    l0 = ctx.lib.rvk_launch_count()
    ms_total = ctx.timed(lambda: step(images), K)
    launches = ctx.lib.rvk_launch_count() - l0
    step_ms = ms_total / K
    value = batch * ctx.world * K / (ms_total / 1e3)
    res = {'metric': metric_name(True), 'value': value, 'unit': 'images/sec', 'ms_per_step': step_ms, 'steps': K, 'warmup': W,
           'batch_per_gpu': batch, 'global_batch': batch * ctx.world, 'gpu_launches': int(launches),
           'optimizer': 'FusedAdamW (rvk_optimizer_*: unscale + inf check + global-norm clip + AdamW + bf16 weight shadows)'
           if fused_tail else 'torch.optim.AdamW(fused=True) + clip_grad_norm_',
           'config': workload_config(True, batch, ctx.world)}
    # phases (device time between events on the launch stream, max over ranks)
    ev['on'] = True
    ctx.barrier()
    for _ in range(K):
        step(images)
    ctx.barrier()
    ev['on'] = False
    marks = ev.pop('marks')
    ph = {'forward_backward': 0.0, 'allreduce_exposed': 0.0, 'clip_optimizer': 0.0}
    for i in range(0, len(marks), 4):
        (_, a), (_, b), (_, c), (_, d) = marks[i:i + 4]
        ph['forward_backward'] += a.elapsed_time(b)
        ph['allreduce_exposed'] += b.elapsed_time(c)
        ph['clip_optimizer'] += c.elapsed_time(d)
    res['phases_ms'] = {k: ctx.max_over_ranks(v / K) for k, v in ph.items()}
    res['allreduce'] = {'ms_exposed': res['phases_ms']['allreduce_exposed'], 'bytes': sum(p.numel() for p in params) * 4,
                        'collectives_per_step': ev.get('collectives', 0), 'overlapped_with_backward': bool(overlap),
                        'buckets': rdist.overlap_state()['buckets'] if overlap else 0,
                        'scale_1_over_world': 'folded into the optimizer kernel' if not average else 'separate mul_'}
    if ctx.world > 1:      # the same payload reduced on its own (nothing to hide behind): hidden = standalone - exposed
        buf = torch.zeros(sum(p.numel() for p in params), device=ctx.dev)
        for _ in range(3):
            dist.all_reduce(buf)
        ms_ar = ctx.timed(lambda: dist.all_reduce(buf), 10) / 10
        res['allreduce']['ms_standalone'] = ms_ar
        res['allreduce']['ms_hidden'] = max(0.0, ms_ar - res['allreduce']['ms_exposed'])
        res['allreduce']['algbw_gbs'] = buf.numel() * 4 / (ms_ar * 1e-3) / 1e9
        del buf
        # A/B: the other all-reduce schedule, same process, same data
        other = not bool(overlap)
        if other:
            rdist.enable_overlap(int(os.environ.get('RVK_DP_BUCKETS', '3')))
        else:
            rdist.disable_overlap()
        for _ in range(3):
            step(images)
        ms_o = ctx.timed(lambda: step(images), K)
        res['allreduce']['overlap_ab'] = {'overlapped_with_backward': other, 'value': batch * ctx.world * K / (ms_o / 1e3),
                                          'ms_per_step': ms_o / K, 'collectives_per_step': ev.get('collectives', 0)}
        if overlap:
            rdist.enable_overlap(int(os.environ.get('RVK_DP_BUCKETS', '3')))
        else:
            rdist.disable_overlap()
    if with_variants:
        for _ in range(2):
            step(images, cutmix=True)
        ms_c = ctx.timed(lambda: step(images, cutmix=True), K)
        res['cutmix'] = {'value': batch * ctx.world * K / (ms_c / 1e3), 'ms_per_step': ms_c / K,
                         'note': 'two JointLoss evaluations (labels_a, labels_b) blended with lam = 0.7, trainer.py:104-111'}
        res['e2e'] = ctx.e2e(step, host_images, K, batch, 4)
    kernels = ctx.kernel_table(lambda: step(images), K, step_ms)
    res['kernels'] = kernels
    pk = ctx.pk
    tc = [e for n, e in kernels.items() if n in TENSOR_BOUND]
    tc_ms = sum(e['ms_per_step'] for e in tc)
    tc_fl = sum(e['gflop_per_launch'] * e['launches_per_step'] for e in tc) * 1e9
    traffic, traffic_src = traffic_entry('train', batch)
    # dominant single kernel of the train step: the weight-gradient GEMM (one symbol, 49 launches per step); the four
    # gemm_nt epilogue variants together take more time but are four different kernels (see `kernels`)
    dom = kernels.get('gemm_tn_kernel')
    res['roofline'] = {'bound': 'tensor', 'achieved': dom['tflops'] if dom else None, 'peak': pk['tflops_sustained'],
                       'unit': 'TFLOP/s', 'frac': dom['frac_tensor'] if dom else None,
                       'kernel': 'gemm_tn_kernel<192,4> (weight gradients dW = A^T B over the token rows, both operands MN-major, split over '
                                 'the rows, red.global.add epilogue; algorithmic FLOPs 2*M*P*Q per launch)',
                       'us_per_launch': dom['us_per_launch'] if dom else None, 'gflop_per_launch': dom['gflop_per_launch'] if dom else None,
                       'share_of_step': dom['share_of_step'] if dom else None, 'launches_timed': dom['launches'] if dom else 0,
                       'traffic': traffic, 'traffic_source': traffic_src, 'peak_source': pk['source'] + ' sustained bf16',
                       'all_tcgen05_kernels': {'tflops': tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else None, 'ms_per_step': tc_ms,
                                               'frac': tc_fl / (tc_ms * 1e-3) / 1e12 / pk['tflops_sustained'] if tc_ms > 0 else None},
                       'whole_step_tflops': value / ctx.world * TRAIN_FLOP_PER_IMG / 1e12,
                       'whole_step_frac': value / ctx.world * TRAIN_FLOP_PER_IMG / 1e12 / pk['tflops_sustained']}
    if with_variants:
        # epochs 1-5 of the reference schedule (trainer.py:244-246): backbone frozen, only the 181 978 head parameters train
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):      # the module prints "Backbone frozen" like the reference (backbone.py:30);
            model.freeze_backbone()                       # stdout carries the ONE JSON line only
        for p in params:
            p.grad = None
        opt, fused_tail = build_optimizer(model, fused_tail)
        if ctx.world > 1 and fused_tail:
            opt.grad_mult = 1.0 / ctx.world
        params = [p for p in model.parameters()]
        for _ in range(3):
            step(images)
        ms_f = ctx.timed(lambda: step(images), K)
        res['frozen_backbone'] = {'value': batch * ctx.world * K / (ms_f / 1e3), 'ms_per_step': ms_f / K,
                                  'trainable_params': sum(p.numel() for p in params if p.requires_grad)}
    rdist.disable_overlap()
    del model, opt, images, host_images
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------ configs[2]: KAN microbenchmark
def bench_kan(ctx, K, W, batch):
    """KANSeverityModule([192,64,1]) (and the production [192,64,16,1]), grid=5 k=3, forward + backward (dx, dW, dWl, db of
    every layer).  Algorithmic bytes / FLOPs per fwd+bwd from SURVEY.md section 8(d): 152 MB, 38.7 GFLOP at batch 65536."""
    torch = ctx.torch
    from rovitkan_b200.models.kan import KANSeverityModule
    out, kernels = {}, None
    for name, dims in (('192-64-1', [192, 64, 1]), ('192-64-16-1', [192, 64, 16, 1])):
        torch.manual_seed(0)
        m = KANSeverityModule(dims).to(ctx.dev)
        x = torch.randn(batch, 192, generator=torch.Generator().manual_seed(0)).to(ctx.dev).requires_grad_(True)
        gy = torch.randn(batch, 1, generator=torch.Generator().manual_seed(1)).to(ctx.dev)

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
            l0 = ctx.lib.rvk_launch_count()
            for _ in range(K):
                ctx.flush_l2()                      # x is 50 MB < 126 MB L2
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            res[tag] = {'ms': ctx.max_over_ranks(tot / K), 'launches': (ctx.lib.rvk_launch_count() - l0) // K}
        # the same six launches replayed from a CUDA graph: the Python autograd round trip between the launches (about as
        # long as the kernels themselves at this size) drops out; values are identical (same kernels, same buffers)
        try:
            side = torch.cuda.Stream(device=ctx.dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    fwdbwd()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                fwdbwd()
            for _ in range(W):
                graph.replay()
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(K):
                ctx.flush_l2()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                graph.replay()
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            res['fwd_bwd_graph'] = {'ms': ctx.max_over_ranks(tot / K), 'launches': res['fwd_bwd']['launches']}
            del graph
        except Exception as e:                     # capture is an extra; the eager number above stands on its own
            res['fwd_bwd_graph'] = {'error': repr(e)[:200]}
            torch.cuda.synchronize()
        out[name] = res
        if name == '192-64-1':
            kernels = ctx.kernel_table(fwdbwd, K, min(res['fwd_bwd']['ms'], res['fwd_bwd_graph'].get('ms', 1e9)))
        del m, x, gy
    pk = ctx.pk
    r = out['192-64-1']
    ms_eager = r['fwd_bwd']['ms']
    ms_graph = r.get('fwd_bwd_graph', {}).get('ms')
    graphed = ms_graph is not None and ms_graph < ms_eager
    ms = ms_graph if graphed else ms_eager
    gbs = KAN_BYTES_PER_SAMPLE * batch / (ms * 1e-3) / 1e9
    traffic, traffic_src = traffic_entry('kan', batch)
    return {'metric': 'samples/sec KANSeverityModule([192,64,1]) forward+backward', 'value': batch * ctx.world / (ms * 1e-3),
            'unit': 'samples/sec', 'ms_per_step': ms, 'ms_per_step_eager_launch': ms_eager, 'fwd_ms': r['fwd']['ms'], 'steps': K,
            'warmup': W, 'batch_per_gpu': batch, 'dtype': 'f32', 'gpu_launches': int(r['fwd_bwd']['launches']) * K,
            'config': {'workload': 'KANLayer microbench: 192->64->1 spline head, grid=5 k=3, batch %d fwd+bwd' % batch,
                       'l2': 'L2 flushed (256 MB memset) between timed iterations',
                       'launch': ('torch.cuda.graph replay of module(x) + backward (same kernels as the eager call; '
                                  '`ms_per_step_eager_launch` is the Python-driven call)') if graphed else 'eager module call'},
            'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': gbs / pk['hbm_gbs'],
                         'traffic': traffic, 'traffic_source': traffic_src,
                         'kernel': 'whole fwd+bwd of the [192,64,1] stack (algorithmic 152 MB / 38.7 GFLOP at batch 65536); per kernel family: `kernels`',
                         'dense_equivalent_tflops': KAN_FLOP_PER_SAMPLE * batch / (ms * 1e-3) / 1e12},
            'kernels': kernels, 'detail': out}


# ------------------------------------------------------------------------------------------ configs[4]: throughput sweep
def bench_sweep(ctx, K, W):
    """Forward throughput of DeiT-Tiny backbone + heads + KAN at batch 64..8192 per GPU, each rank on its own shard (no
    collective), against the tensor roofline (2.507 GFLOP per image); L2 flushed before every timed step."""
    torch, dist = ctx.torch, ctx.dist
    from rovitkan_b200.models import RoViTKAN
    torch.manual_seed(0)
    model = RoViTKAN(pretrained=False).to(ctx.dev).eval()
    rows = []
    for batch in SWEEP_BATCHES:
        g = torch.Generator(device=ctx.dev).manual_seed(1000 + ctx.rank)
        images = torch.randn(batch, 3, 224, 224, generator=g, device=ctx.dev)
        with torch.no_grad():
            for _ in range(W):
                model(images)
            tot = 0.0
            for _ in range(K):
                ctx.flush_l2()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ctx.barrier()
                e0.record()
                model(images)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
        ms = ctx.max_over_ranks(tot / K)
        ips = batch * ctx.world / (ms / 1e3)
        rows.append({'batch_per_gpu': batch, 'ms_per_step': ms, 'images_per_sec': ips,
                     'tensor_roofline_frac': ips / ctx.world * FWD_FLOP_PER_IMG / 1e12 / ctx.pk['tflops_sustained']})
        del images
    del model
    torch.cuda.empty_cache()
    best = max(rows, key=lambda r: r['images_per_sec'])
    return {'metric': 'images/sec (224^2, device-timed) RoViT-KAN inference forward, batch sweep', 'unit': 'images/sec',
            'steps': K, 'warmup': W, 'value': best['images_per_sec'], 'best_batch_per_gpu': best['batch_per_gpu'],
            'config': {'workload': 'encoder-throughput sweep, batch 64-8192 per GPU', 'parallelism': f'dp{ctx.world}',
                       'l2': 'L2 flushed (256 MB memset) before every timed step'},
            'peak_tflops': ctx.pk['tflops_sustained'], 'rows': rows}


# ------------------------------------------------------------------------------------------ PyTorch eager on the same GPU
def eager_trunk_sdpa(sd, x, prefix='backbone.model.'):
    """timm's deit_tiny_patch16_224 forward as timm >= 0.9 runs it on a GPU: library kernels only (cuDNN conv, cuBLAS linears,
    F.scaled_dot_product_attention = the flash / mem-efficient SDPA backend, ATen LayerNorm / GELU)."""
    import torch
    import torch.nn.functional as F
    g = lambda k: sd[prefix + k]
    t = F.conv2d(x, g('patch_embed.proj.weight').to(x.dtype), g('patch_embed.proj.bias').to(x.dtype), stride=16).flatten(2).transpose(1, 2)
    t = torch.cat([g('cls_token').to(t.dtype).expand(t.shape[0], -1, -1), t], dim=1) + g('pos_embed').to(t.dtype)
    B, N, D = t.shape
    for i in range(12):
        b = f'blocks.{i}.'
        h = F.layer_norm(t, (D,), g(b + 'norm1.weight'), g(b + 'norm1.bias'), 1e-6)
        qkv = F.linear(h, g(b + 'attn.qkv.weight'), g(b + 'attn.qkv.bias')).reshape(B, N, 3, 3, 64).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2]).transpose(1, 2).reshape(B, N, D)
        t = t + F.linear(a, g(b + 'attn.proj.weight'), g(b + 'attn.proj.bias'))
        h = F.layer_norm(t, (D,), g(b + 'norm2.weight'), g(b + 'norm2.bias'), 1e-6)
        t = t + F.linear(F.gelu(F.linear(h, g(b + 'mlp.fc1.weight'), g(b + 'mlp.fc1.bias'))), g(b + 'mlp.fc2.weight'), g(b + 'mlp.fc2.bias'))
    return F.layer_norm(t, (D,), g('norm.weight'), g('norm.bias'), 1e-6)[:, 0]


def gpu_eager_baseline(ctx, batch_infer=1024, batch_train=256, steps=5):
    """The "existing Blackwell kernels" bar (SURVEY.md section 2 / BASELINE.md section 3): the oracle restatement of the
    reference modules run by PyTorch eager on this B200 (cuBLAS / SDPA / ATen kernels; the KAN contraction vectorised as an
    einsum -- the reference's own Python loop would take ~0.4 s per batch), fp32 with TF32 off and bf16 autocast."""
    torch = ctx.torch
    from oracle import losses as olosses
    from oracle import model as omodel
    from oracle import vit as ovit
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    sd = {k: v.to(ctx.dev) for k, v in omodel.random_state_dict(0).items()}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch_infer, 3, 224, 224, generator=g).to(ctx.dev)
    out = {'note': 'oracle/vit.py + vectorised KAN + heads through PyTorch eager on the same GPU, CUDA-event timed', 'steps': steps}

    def fwd(autocast, sdpa=False):
        trunk = eager_trunk_sdpa if sdpa else (lambda s, xx: ovit.forward_functional(s, xx, prefix='backbone.model.'))
        with torch.no_grad():
            if autocast:
                with torch.autocast('cuda', dtype=torch.bfloat16):
                    f = trunk(sd, x).float()
            else:
                f = trunk(sd, x)
            return omodel.heads_forward(sd, f, 4)['kan_severity']

    def time_it(fn, n):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    for tag, ac, sdpa in (('infer_fp32_no_tf32', False, False), ('infer_bf16_autocast', True, False),
                          ('infer_bf16_autocast_sdpa', True, True)):
        ms = time_it(lambda: fwd(ac, sdpa), steps)
        out[tag] = {'batch': batch_infer, 'ms_per_step': ms, 'images_per_sec': batch_infer / (ms / 1e3)}
    with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):       # the SDPA restatement is the same function
        a, b = eager_trunk_sdpa(sd, x[:8]).float(), ovit.forward_functional(sd, x[:8], prefix='backbone.model.').float()
    out['sdpa_trunk_vs_oracle_rel_l2'] = float((a - b).norm() / b.norm())
    del x
    sdt = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'knots' not in k else v) for k, v in sd.items()}
    xt = torch.randn(batch_train, 3, 224, 224, generator=g).to(ctx.dev)
    y = torch.randint(0, 4, (batch_train,), generator=g).to(ctx.dev)

    def train_step():
        for v in sdt.values():
            if v.requires_grad:
                v.grad = None
        with torch.autocast('cuda', dtype=torch.bfloat16):
            f = eager_trunk_sdpa(sdt, xt).float()
        olosses.joint(omodel.heads_forward(sdt, f, 4), y, y, 4)['total_loss'].backward()
    ms = time_it(train_step, steps)
    out['train_bf16_autocast_sdpa_fwd_bwd'] = {'batch': batch_train, 'ms_per_step': ms, 'images_per_sec': batch_train / (ms / 1e3),
                                          'note': 'forward + joint loss + backward, no optimizer'}
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
    del sd, sdt, xt
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------ drivers
def headline(ctx, args, res, train):
    K, W = args.steps, max(3, args.warmup)
    batch = res['batch'] if not train else res['batch_per_gpu']
    line = {'metric': metric_name(train), 'value': res['value'], 'unit': 'images/sec', 'n_gpus': ctx.world, 'steps': K,
            'warmup': W, 'ms_per_step': res['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'bf16', 'data': 'synthetic', 'config': workload_config(train, batch, ctx.world), 'e2e': res.get('e2e'),
            'gpu_launches': res['gpu_launches'], 'clocks': res.get('clocks'), 'roofline': res['roofline']}
    for k in ('e2e_bf16_input', 'e2e_uint8_input', 'phases_ms', 'allreduce', 'cutmix', 'frozen_backbone', 'optimizer'):
        if k in res:
            line[k] = res[k]
    if train:
        line['kernels'] = res['kernels']
    return line


def guarded_sub(fn, world):
    """A sub-benchmark (configs[2..4]) next to the headline.  On one GPU its failure is recorded in its own sub-object and the
    headline line is still printed; with several ranks an exception stays fatal (the other ranks would wait in a barrier
    or collective the failed rank never reaches, and torchrun tears the job down at once instead)."""
    if world > 1:
        return fn()
    try:
        return fn()
    except Exception as e:
        import traceback
        traceback.print_exc(file=sys.stderr)
        return {'error': repr(e)}


def run_ours(args, rank, local_rank, world):
    ctx = Ctx(rank, local_rank, world)
    K, W = args.steps, max(3, args.warmup)
    mode = args.mode
    line = None
    if mode in ('all', 'infer'):
        res = bench_infer(ctx, K, W, args.batch or 1024)
        line = headline(ctx, args, res, False)
        if mode == 'all' and not args.no_subs:
            Ks = max(5, min(K, 20))
            for key, fn in (('train', lambda: bench_train(ctx, Ks, W, 256)), ('kan', lambda: bench_kan(ctx, Ks, W, 65536)),
                            ('sweep', lambda: bench_sweep(ctx, max(3, min(K, 8)), 3))):
                line[key] = guarded_sub(fn, world)
            if rank == 0 and world == 1:
                try:
                    line['gpu_eager_baseline'] = gpu_eager_baseline(ctx)
                    fast = line['gpu_eager_baseline']['infer_bf16_autocast_sdpa']['images_per_sec']
                    line['gpu_eager_baseline']['ours_over_eager_bf16_infer'] = line['value'] / fast
                    line['gpu_eager_baseline']['ours_over_eager_bf16_train'] = (
                        line['train']['value'] / line['gpu_eager_baseline']['train_bf16_autocast_sdpa_fwd_bwd']['images_per_sec'])
                except Exception as e:      # a baseline leg must never take the measurement down
                    line['gpu_eager_baseline'] = {'error': repr(e)}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            r = cpu_reference(steps=args.cpu_steps, warmup=1)
            line['cpu_baseline'] = {k: r[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
            if mode == 'all' and not args.no_subs:
                t = cpu_reference_train(steps=1, warmup=1)
                line['train']['cpu_baseline'] = {k: t[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
                line['kan']['cpu_baseline'] = cpu_reference_kan()
    elif mode == 'train':
        res = bench_train(ctx, K, W, args.batch or 256)
        line = headline(ctx, args, res, True)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            t = cpu_reference_train(steps=2, warmup=1)
            line['cpu_baseline'] = {k: t[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')}
    elif mode == 'kan':
        res = bench_kan(ctx, K, W, args.batch or 65536)
        line = dict(res, n_gpus=world, higher_is_better=True, scaling='weak', vs_baseline=None, data='synthetic')
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_reference_kan(batches=(32, 1024, 4096))
    elif mode == 'sweep':
        res = bench_sweep(ctx, K, W)
        line = dict(res, n_gpus=world, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='bf16', data='synthetic')
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mode', default='all', choices=['all', 'infer', 'train', 'kan', 'sweep'])
    ap.add_argument('--batch', type=int, default=0)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-subs', action='store_true', help='mode all: skip the train / kan / sweep / eager sub-benchmarks')
    ap.add_argument('--cpu-steps', type=int, default=40, help='batch-32 reference forwards timed for cpu_baseline (~10-20 s)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == '__main__':
    main()
