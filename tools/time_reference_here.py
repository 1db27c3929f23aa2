"""Build container only (needs /root/reference; the GPU box has none): time the REFERENCE's own modules on the host cores
next to the oracle port that `bench.py --impl reference` / `cpu_baseline` run on the GPU box (kind "port"), same weights, same
inputs, same thread count -- the evidence that the port's CPU time stands for the reference's.

    python tools/time_reference_here.py  >  profiles/r2_cpu_reference_vs_port.json

The reference's `models/rovit_kan.py` is imported unmodified; the absent third-party `timm` is replaced by the oracle's
restatement of deit_tiny_patch16_224 (oracle/vit.py), exactly as tests/golden/make_golden.py does.  Workloads = BASELINE.json
configs[0]: eval forward at batch 32 (stage 4, fp32) and the stage-4 forward + JointLoss + backward at batch 32.
"""

import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get('ROVIT_REFERENCE', '/root/reference')
sys.path.insert(0, ROOT)
from oracle import losses as olosses  # noqa: E402
from oracle import model as omodel  # noqa: E402
from oracle import vit as ovit  # noqa: E402

shim = types.ModuleType('timm')
shim.create_model = ovit.create_model
sys.modules['timm'] = shim
sys.path.insert(0, REF)
from models.rovit_kan import RoViTKAN  # noqa: E402   (the reference's)
from training.losses import JointLoss  # noqa: E402   (the reference's)

assert os.path.abspath(sys.modules['models.rovit_kan'].__file__).startswith(os.path.abspath(REF))


def timed_pair(fn_ref, fn_port, rounds):
    """The two callables alternate (the build container shares its cores: back-to-back blocks of one implementation pick up
    whatever else runs at the time); returns (median, min) seconds of each."""
    fn_ref(); fn_port()
    tr, tp = [], []
    for _ in range(rounds):
        for fn, acc in ((fn_ref, tr), (fn_port, tp)):
            t0 = time.perf_counter()
            fn()
            acc.append(time.perf_counter() - t0)
    tr.sort(); tp.sort()
    return (tr[len(tr) // 2], tr[0]), (tp[len(tp) // 2], tp[0])


def main():
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = 32
    sd = omodel.random_state_dict(0)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, 3, 224, 224, generator=g)
    y = torch.randint(0, 4, (batch,), generator=g)

    ref = RoViTKAN(pretrained=False, dropout=0.0)
    ref.load_state_dict(sd)
    ref.curriculum_stage = 4
    ref.eval()
    with torch.no_grad():
        a = ref(x)
        b = omodel.forward(sd, x, stage=4, kan_loop=True)
        dev = {k: float((a[k] - b[k]).abs().max()) for k in ('features', 'cls_logits', 'ordinal_logits', 'mu', 'log_var', 'kan_severity')}
        fwd_ref, fwd_port = timed_pair(lambda: ref(x), lambda: omodel.forward(sd, x, stage=4, kan_loop=True), 7)

    ref.train()
    loss_fn = JointLoss()

    def ref_step():
        ref.zero_grad(set_to_none=True)
        loss_fn(ref(x), y, y, 4)['total_loss'].backward()

    sdt = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'knots' not in k else v) for k, v in sd.items()}

    def port_step():
        for v in sdt.values():
            if v.requires_grad:
                v.grad = None
        olosses.joint(omodel.forward(sdt, x, stage=4, kan_loop=True), y, y, 4)['total_loss'].backward()

    step_ref, step_port = timed_pair(ref_step, port_step, 5)

    def leg(r, p):
        return {'reference_s_median': r[0], 'port_s_median': p[0], 'reference_s_min': r[1], 'port_s_min': p[1],
                'reference_img_s': batch / r[0], 'port_img_s': batch / p[0], 'port_over_reference_time_median': p[0] / r[0],
                'port_over_reference_time_min': p[1] / r[1]}
    out = {'where': 'build container (no GPU)', 'cores': cores, 'torch': torch.__version__, 'batch': batch,
           'reference_modules': 'models/rovit_kan.py + training/losses.py of /root/reference, timm -> oracle/vit.py::create_model',
           'max_abs_output_difference_reference_vs_port': dev,
           'eval_forward': leg(fwd_ref, fwd_port), 'train_fwd_loss_bwd': leg(step_ref, step_port)}
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
