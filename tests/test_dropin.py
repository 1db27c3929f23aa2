"""Drop-in boundary (SURVEY.md section 8b "Form", N1): the reference's scripts import `models.*`, `training.*`, `data.*` by
those names after inserting their own root at sys.path[0]; rovitkan_b200.install() must serve the sm_100a modules under
those names, leave `training.optimizer` / `training.trainer` to the reference, and supply the git-ignored `data` package."""

import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = '/root/reference'
PLOT_STUBS = ("import sys\nfrom unittest import mock\n"
              "for _n in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.patches', 'matplotlib.cm', 'matplotlib.colors', "
              "'matplotlib.gridspec', 'seaborn'):\n"
              "    try:\n        __import__(_n)\n    except ImportError:\n        sys.modules[_n] = mock.MagicMock()\n")


def _run(code, cwd, timeout=300, env_extra=None):
    env = dict(os.environ, PYTHONPATH=ROOT, ROVITKAN_SYNTH_PER_CLASS='4', ROVITKAN_DATA_WORKERS='0', CUDA_VISIBLE_DEVICES='',
               ROVITKAN_PRETRAINED='random',      # the reference's config asks for pretrained=True; no weights on this box
               ROVITKAN_SYNTH_DIR=os.path.join(str(cwd), 'synthetic_images'))
    env.update(env_extra or {})
    return subprocess.run([sys.executable, '-c', code], cwd=str(cwd), env=env, capture_output=True, text=True, timeout=timeout)


def _fake_reference(tmp_path):
    """A tree shaped like the reference (scripts/ insert the root at sys.path[0]; training/ has optimizer.py AND a
    losses.py that must lose against ours; models/ must lose entirely; no data/)."""
    root = tmp_path / 'ref'
    for d in ('scripts', 'training', 'models', 'configs'):
        (root / d).mkdir(parents=True)
        (root / d / '__init__.py').write_text('')
    (root / 'training' / 'optimizer.py').write_text('WHO = "reference"\n')
    (root / 'training' / 'losses.py').write_text('WHO = "reference"\nraise RuntimeError("reference losses imported")\n')
    (root / 'models' / 'rovit_kan.py').write_text('raise RuntimeError("reference models imported")\n')
    (root / 'configs' / 'config.py').write_text('WHO = "reference"\n')
    (root / 'scripts' / 'run.py').write_text(textwrap.dedent('''
        import sys
        from pathlib import Path
        sys.path.insert(0, str(Path(__file__).parent.parent))
        from configs.config import WHO as cfg_who
        from data.dataset import create_dataloaders, RoseLeafDataset
        from data.transforms import augmented_transforms, original_transforms, cutmix_or_mixup
        from models.rovit_kan import RoViTKAN
        from models.kan import KANLayer
        from training.losses import JointLoss
        from training.optimizer import WHO as opt_who
        import training, models, data, rovitkan_b200.models.rovit_kan as ours
        assert cfg_who == opt_who == 'reference'
        assert RoViTKAN is ours.RoViTKAN and JointLoss.__module__ == 'rovitkan_b200.training.losses'
        assert data.__name__ == 'rovitkan_b200.data' and models.__name__ == 'rovitkan_b200.models'
        assert sys.argv[1:] == ['--flag', '7'], sys.argv
        print('DROPIN_OK', training.__path__)
    '''))
    return root


def test_hook_serves_ours_and_lets_the_reference_through(tmp_path):
    root = _fake_reference(tmp_path)
    r = subprocess.run([sys.executable, '-m', 'rovitkan_b200.launch', str(root / 'scripts' / 'run.py'), '--flag', '7'],
                       cwd=str(tmp_path), env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and 'DROPIN_OK' in r.stdout, r.stdout + r.stderr
    assert str(root / 'training') in r.stdout


def test_real_data_package_wins_over_the_synthetic_one(tmp_path):
    root = _fake_reference(tmp_path)
    (root / 'data').mkdir()
    (root / 'data' / '__init__.py').write_text('')
    (root / 'data' / 'dataset.py').write_text('WHO = "user data"\n')
    code = textwrap.dedent(f'''
        import sys, rovitkan_b200
        rovitkan_b200.install()
        sys.path.insert(0, {str(root)!r})
        from data.dataset import WHO
        assert WHO == 'user data'
        rovitkan_b200.uninstall(); rovitkan_b200.install(synthetic_data=True)
        for n in [m for m in sys.modules if m == 'data' or m.startswith('data.')]: del sys.modules[n]
        from data.dataset import RoseLeafDataset
        assert RoseLeafDataset.__module__ == 'rovitkan_b200.data.dataset'
        print('OK')
    ''')
    r = _run(code, tmp_path)
    assert r.returncode == 0 and 'OK' in r.stdout, r.stdout + r.stderr


def test_install_after_reference_import_is_refused(tmp_path):
    root = _fake_reference(tmp_path)
    code = textwrap.dedent(f'''
        import sys
        sys.path.insert(0, {str(root)!r})
        import training.optimizer
        import rovitkan_b200
        try:
            rovitkan_b200.install()
        except RuntimeError as e:
            print('REFUSED', e)
    ''')
    r = _run(code, tmp_path)
    assert 'REFUSED' in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='the reference tree is not on this machine')
def test_reference_modules_resolve_through_the_hook(tmp_path):
    code = PLOT_STUBS + textwrap.dedent(f'''
        import rovitkan_b200
        rovitkan_b200.install()
        sys.path.insert(0, {REFERENCE!r})
        import training.optimizer, training.trainer, training.losses, models.rovit_kan, models.heads, data.transforms
        import evaluation.evaluator, configs.config
        assert training.optimizer.__file__.startswith({REFERENCE!r}) and training.trainer.__file__.startswith({REFERENCE!r})
        assert evaluation.evaluator.__file__.startswith({REFERENCE!r})
        assert training.losses.__name__ == 'rovitkan_b200.training.losses'
        assert training.trainer.cutmix_or_mixup.__module__ == 'rovitkan_b200.data.transforms'
        from rovitkan_b200.models import RoViTKAN
        assert models.rovit_kan.RoViTKAN is RoViTKAN
        cfg = configs.config.get_config()
        m = RoViTKAN(cfg)                                                  # RoViTKAN(config), ablation.py:264
        opt = training.optimizer.build_optimizer(m, cfg)                   # the reference's two LR groups on our parameters
        assert [len(g['params']) for g in opt.param_groups] == [150, 23]
        print('OK')
    ''')
    r = _run(code, tmp_path)
    assert r.returncode == 0 and 'OK' in r.stdout, r.stdout + r.stderr


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason='the reference tree is not on this machine')
@pytest.mark.parametrize('script', ['train.py', 'evaluate.py', 'run_ablation.py'])
def test_unmodified_reference_scripts_run_up_to_the_cuda_boundary(tmp_path, script):
    """No GPU here and no reference tree on the GPU box, so this is as far as the unmodified scripts can be driven: every
    import, the synthetic dataloaders, RoViTKAN(embed_dim=...), build_optimizer, JointLoss, Trainer(...) and trainer.fit() up
    to the first trunk call, which must raise OUR no-CPU-fallback error (a reference import would run on CPU instead)."""
    args = ['--data_root', str(tmp_path / 'nodata')]
    if script == 'train.py':
        args += ['--output_dir', str(tmp_path / 'out')]
    elif script == 'run_ablation.py':            # its own flag spelling (run_ablation.py:45-106); 5 epochs x 7 experiments in --fast
        args = ['--data-root', str(tmp_path / 'nodata'), '--output-dir', str(tmp_path / 'out'), '--fast', '--num-workers', '0']
    else:
        from rovitkan_b200.models import RoViTKAN
        ck = tmp_path / 'ck.pth'
        torch.save({'model_state_dict': RoViTKAN(pretrained=False).state_dict(), 'epoch': 1}, ck)
        args += ['--checkpoint', str(ck), '--batch_size', '4']
    code = PLOT_STUBS + textwrap.dedent(f'''
        from rovitkan_b200 import launch
        try:
            launch.main([{os.path.join(REFERENCE, 'scripts', script)!r}] + {args!r})
        except RuntimeError as e:
            print('BOUNDARY', e)
    ''')
    r = _run(code, tmp_path, timeout=600)
    assert 'BOUNDARY' in r.stdout and 'no CPU fallback' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
    if script == 'train.py':
        assert 'backbone: 5,524,416' in r.stdout and 'Backbone frozen' in r.stdout and 'Epoch 1/' in r.stdout
    if script == 'run_ablation.py':
        assert 'Experiment: full_model' in r.stdout and 'Backbone frozen' in r.stdout and 'Epoch 1/' in r.stdout


# ----------------------------------------------------------------------------------------- the data package itself
def test_synthetic_dataloaders_have_the_surface_the_scripts_use(tmp_path, monkeypatch):
    monkeypatch.setenv('ROVITKAN_SYNTH_PER_CLASS', '5')
    monkeypatch.setenv('ROVITKAN_DATA_WORKERS', '0')
    monkeypatch.setenv('ROVITKAN_SYNTH_DIR', str(tmp_path / 'synthetic_images'))
    from rovitkan_b200.data.dataset import DEFAULT_CLASSES, RoseLeafDataset, create_dataloaders
    from rovitkan_b200.data.transforms import augmented_transforms, inference_transforms, original_transforms
    sev = {c: i for i, c in enumerate(DEFAULT_CLASSES)}
    tr, va, te = create_dataloaders(tmp_path / 'a', tmp_path / 'o', DEFAULT_CLASSES, sev, augmented_transforms(),
                                    original_transforms(), batch_size=4, train_val_split=0.8, num_workers=0, seed=1)
    base = tr.dataset.dataset                                   # scripts/train.py:110 unwraps the Subset
    assert isinstance(base, RoseLeafDataset) and len(base) == 20 and len(tr.dataset) == 16 and len(va.dataset) == 4
    assert torch.allclose(base.get_class_weights(), torch.ones(4))
    assert set(base.samples[0]) == {'path', 'class_idx', 'severity'} and base.classes == DEFAULT_CLASSES
    x, y, s = next(iter(tr))
    assert x.shape == (4, 3, 224, 224) and x.dtype == torch.float32 and y.dtype == torch.int64 and torch.equal(y, s)
    x2, _, _ = next(iter(te))
    assert x2.shape == (4, 3, 224, 224)
    ds = RoseLeafDataset(tmp_path / 'o', DEFAULT_CLASSES, sev, transform=inference_transforms(), mode='original')
    a, b = ds[3][0], ds[3][0]
    assert torch.equal(a, b)                                    # deterministic without augmentation
    # scripts/run_ablation.py:34-42 (TransformSubset) and run_baselines.py re-open samples[i]['path'] with PIL themselves to swap
    # the transform of a Subset: synthetic samples are therefore real files, and both routes see the same pixels
    from PIL import Image
    tf = inference_transforms()
    for i in (0, 7, len(ds) - 1):
        assert os.path.isfile(ds.samples[i]['path'])
        assert torch.equal(tf(Image.open(ds.samples[i]['path']).convert('RGB')), ds[i][0])
    assert ds.samples[0]['path'] != base.samples[0]['path']     # the 'original' (test) images are not the training images


def test_image_folders_are_read_when_present(tmp_path):
    from PIL import Image
    from rovitkan_b200.data.dataset import RoseLeafDataset
    from rovitkan_b200.data.transforms import original_transforms
    for ci, c in enumerate(('A', 'B')):
        (tmp_path / c).mkdir()
        for i in range(2 + ci):
            Image.fromarray(np.full((40, 60, 3), 50 * (ci + 1), np.uint8)).save(tmp_path / c / f'{i}.png')
    ds = RoseLeafDataset(tmp_path, ['A', 'B'], {'A': 0, 'B': 3}, transform=original_transforms())
    assert not ds.synthetic and len(ds) == 5 and ds[4][1:] == (1, 3)
    assert ds[0][0].shape == (3, 224, 224)
    assert torch.allclose(ds.get_class_weights(), torch.tensor([5 / 4, 5 / 6]))


def test_cutmix_or_mixup_contract():
    from rovitkan_b200.data.transforms import cutmix_or_mixup
    np.random.seed(0)
    torch.manual_seed(0)
    x = torch.arange(6, dtype=torch.float32).view(6, 1, 1, 1).expand(6, 3, 32, 32).contiguous()
    y = torch.arange(6)
    seen = set()
    for _ in range(20):
        xm, ya, yb, lam = cutmix_or_mixup(x, y, use_cutmix=True, use_mixup=True, cutmix_alpha=1.0, mixup_alpha=0.2)
        assert isinstance(lam, float) and 0.0 <= lam <= 1.0 and torch.equal(ya, y) and sorted(yb.tolist()) == list(range(6))
        # every pixel is a convex mix of image a and image b, and the mean weight of a is lam
        a = y.view(6, 1, 1, 1).float()
        b = yb.view(6, 1, 1, 1).float()
        w = torch.where(a != b, (xm - b) / (a - b).where(a != b, torch.ones(())), torch.full_like(xm, lam))
        assert bool(((w > -1e-5) & (w < 1 + 1e-5)).all())
        sel = (a != b).expand_as(xm)
        if sel.any():
            assert abs(float(w[sel].mean()) - lam) < 1e-4
        seen.add('cutmix' if bool(((w - w.round()).abs() < 1e-6).all()) else 'mixup')
    assert seen == {'cutmix', 'mixup'}
    xm, ya, yb, lam = cutmix_or_mixup(x, y, use_cutmix=False, use_mixup=False)
    assert lam == 1.0 and xm is x
