"""A script written the way the reference's scripts/train.py + scripts/evaluate.py are (same imports BY THE REFERENCE'S
MODULE NAMES, same call sequence as training/trainer.py:54-153 and evaluation/evaluator.py:25-66), run through
`python -m rovitkan_b200.launch` by tests/test_gpu_dropin_flow.py.  The reference tree itself is not available on the GPU
box, so this stands in for it there; tests/test_dropin.py drives the real scripts up to the CUDA boundary on the CPU box."""

import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).parent.parent / 'no_such_reference_root'))     # scripts/train.py:9 does this with its root

import torch
from data.dataset import RoseLeafDataset, create_dataloaders
from data.transforms import augmented_transforms, cutmix_or_mixup, original_transforms
from models.rovit_kan import RoViTKAN
from training.losses import JointLoss
from torch.amp import GradScaler, autocast


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--output_dir', required=True)
    ap.add_argument('--epochs', type=int, default=2)
    ap.add_argument('--batch_size', type=int, default=8)
    args = ap.parse_args()
    torch.manual_seed(42)
    device = torch.device('cuda')
    classes = ['Healthy Leaf', 'Leaf Holes', 'Black Spot', 'Dry Leaf']
    sev = {c: i for i, c in enumerate(classes)}
    train_loader, val_loader, test_loader = create_dataloaders(
        augmented_root=Path('nodata/Augmented Image'), original_root=Path('nodata/Original Image'), class_names=classes,
        severity_map=sev, augmented_transform=augmented_transforms(), original_transform=original_transforms(),
        batch_size=args.batch_size, train_val_split=0.8, num_workers=0, seed=42)
    model = RoViTKAN(embed_dim=192, hidden_dim=128, num_classes=4, kan_layers=[192, 64, 16, 1], kan_num_knots=5, kan_degree=3,
                     dropout=0.3, pretrained=False).to(device)                        # scripts/train.py:88-97
    backbone = [p for n, p in model.named_parameters() if 'backbone' in n]            # training/optimizer.py:12-25
    heads = [p for n, p in model.named_parameters() if 'backbone' not in n]
    opt = torch.optim.AdamW([{'params': backbone, 'lr': 1e-5}, {'params': heads, 'lr': 1e-4}], weight_decay=1e-4)
    weights = train_loader.dataset.dataset.get_class_weights().to(device)             # scripts/train.py:110-111
    loss_fn = JointLoss(lambda_ord=1.0, mu_unc=0.5, nu_kan=0.5, focal_gamma=2.0, focal_alpha=weights, num_classes=4)
    scaler = GradScaler('cuda')
    history = []
    model.freeze_backbone()                                                           # trainer.py:244-246
    for epoch in range(1, args.epochs + 1):
        model.train()
        stage = 4 if epoch > 1 else 2
        model.curriculum_stage = stage
        if epoch == 2:
            model.unfreeze_backbone()                                                 # trainer.py:62-63
        tot, correct, n = 0.0, 0, 0
        for bi, (images, y, s) in enumerate(train_loader):
            if bi >= 2:
                break
            images, y, s = images.to(device), y.to(device), s.to(device)
            images, ya, yb, lam = cutmix_or_mixup(images, y, use_cutmix=True, use_mixup=True, cutmix_alpha=1.0, mixup_alpha=0.2)
            with autocast('cuda'):
                out = model(images)
                la, lb = loss_fn(out, ya, s, stage), loss_fn(out, yb, s, stage)
                losses = {k: lam * la[k] + (1 - lam) * lb[k] for k in la}
                loss = losses['total_loss']
            opt.zero_grad()
            scaler.scale(loss).backward()
            scaler.unscale_(opt)
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            scaler.step(opt)
            scaler.update()
            tot += loss.item()
            correct += out['cls_logits'].max(1)[1].eq(y).sum().item()
            n += y.size(0)
        model.eval()
        vl = 0.0
        with torch.no_grad():
            for images, y, s in val_loader:
                out = model(images.to(device))
                vl += loss_fn(out, y.to(device), s.to(device), stage=4)['total_loss'].item()   # trainer.py:205
        history.append({'epoch': epoch, 'stage': stage, 'train_loss': tot / 2, 'val_loss': vl / len(val_loader), 'acc': correct / n})
    ck = Path(args.output_dir) / 'best_model.pth'
    torch.save({'epoch': args.epochs, 'model_state_dict': model.state_dict(), 'optimizer_state_dict': opt.state_dict(),
                'scaler_state_dict': scaler.state_dict()}, ck)                        # trainer.py:311-325
    # ---- scripts/evaluate.py + evaluator.py:229-253, 25-66
    model2 = RoViTKAN(embed_dim=192, hidden_dim=128, num_classes=4, kan_layers=[192, 64, 16, 1], pretrained=False)
    model2.load_state_dict(torch.load(ck, map_location=device, weights_only=False)['model_state_dict'])
    model2.to(device).eval()
    test = RoseLeafDataset(root_dir=Path('nodata/Original Image'), class_names=classes, severity_map=sev,
                           transform=original_transforms(), mode='original')
    loader = torch.utils.data.DataLoader(test, batch_size=args.batch_size, shuffle=False, num_workers=0)
    preds, sevs, same = [], [], True
    with torch.no_grad():
        for images, y, s in loader:
            o = model2(images.to(device))
            o1 = model(images.to(device))
            same = same and all(torch.equal(o[k], o1[k]) for k in o)
            preds.append(torch.argmax(torch.softmax(o['cls_logits'], 1), 1).cpu())
            sevs.append(o['kan_severity'].squeeze().cpu())
            assert torch.isfinite(torch.exp(0.5 * o['log_var'])).all()
    pred = model2.predict(next(iter(loader))[0].to(device))
    print('FLOW_RESULT ' + json.dumps({'history': history, 'n_test': int(torch.cat(preds).numel()), 'reload_identical': same,
                                       'predict_keys': sorted(pred), 'sev_range': [float(torch.cat(sevs).min()), float(torch.cat(sevs).max())]}))


if __name__ == '__main__':
    main()
