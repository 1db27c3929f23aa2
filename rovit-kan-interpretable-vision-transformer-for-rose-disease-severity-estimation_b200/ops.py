"""torch.autograd glue over the C ABI (include/rovitkan.h).

Every op here allocates its outputs/workspaces with torch, passes raw device pointers plus the
current CUDA stream to librovitkan.so, and raises on CPU tensors: the path has no CPU fallback.
PyTorch is used for device memory, streams and autograd bookkeeping only.
"""

from __future__ import annotations

import ctypes as C
import os

import torch
from torch.amp import custom_bwd, custom_fwd

from . import _lib

# images per pass through the 12 blocks; measured on B200 (tools/sweep_chunk.sh): throughput rises monotonically
# with the chunk (55k img/s at 48 ... 123k at 1024), so the default only bounds the workspace (1.2 GB per 1024 images)
CHUNK_IMAGES = int(os.environ.get('ROVITKAN_CHUNK_IMAGES', '2048'))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t) -> int:
    return 0 if t is None else t.data_ptr()


def nograd(*tensors):
    """Under torch.no_grad() hand the Functions detached tensors: `ctx.needs_input_grad` reflects requires_grad of
    the inputs even when grad mode is off, and would otherwise select the activation-saving training kernels."""
    if torch.is_grad_enabled():
        return tensors
    return tuple(t.detach() if isinstance(t, torch.Tensor) else t for t in tensors)


def require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f'{what}: got a {t.device.type} tensor. rovitkan_b200 runs only on CUDA (sm_100a) through '
            f'librovitkan.so and has no CPU fallback; move the module and its inputs to a B200 device.')


def _f32c(t: torch.Tensor) -> torch.Tensor:
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _host_floats(values):
    return (C.c_float * len(values))(*values)


def _rng_keys():
    """(seed, offset) for the Philox dropout stream, drawn from torch's CPU generator so that
    torch.manual_seed() controls it and no device sync is needed."""
    r = torch.randint(0, 2 ** 62, (2,), dtype=torch.int64)
    return int(r[0]), int(r[1])


# --------------------------------------------------------------------------------------------- KAN
def kan_basis(t: torch.Tensor, knots_host) -> torch.Tensor:
    """BSplineBasis.compute_basis (reference models/kan.py:10-44) for already-normalised inputs."""
    require_cuda(t, 'BSplineBasis.compute_basis')
    tc = _f32c(t)
    out = torch.empty(*tc.shape, 7, device=tc.device, dtype=torch.float32)
    with torch.cuda.device(tc.device):
        _lib.call('rvk_kan_basis', _p(tc), _host_floats(knots_host), len(knots_host), tc.numel(), _p(out), _stream())
    return out


class KanLayerFn(torch.autograd.Function):
    """One fused KAN layer: y = act(Linear(x) + sum_i sum_k N_k(tanh x_i) W[i,:,k]) (kan.py:70-95)."""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, x, spline, lin_w, lin_b, knots_host, act, state=None):
        require_cuda(x, 'KANLayer.forward')
        xc, sw, lw, lb = _f32c(x), _f32c(spline), _f32c(lin_w), _f32c(lin_b)
        batch, n_in = xc.shape
        n_out = lw.shape[0]
        need_bwd = any(ctx.needs_input_grad[:4])
        lib = _lib.load()
        # packed / split weight operands live in a workspace that is reused while the parameters are unchanged
        key = (xc.device, need_bwd, tuple(knots_host), tuple((t.data_ptr(), t._version, t.dtype) for t in (spline, lin_w)))
        prepared = state is not None and state.key == key and state.ws is not None
        if prepared:
            ws = state.ws
        else:
            ws = torch.empty(lib.rvk_kan_layer_workspace_floats(n_in, n_out, int(need_bwd)), device=xc.device,
                             dtype=torch.float32)
            if state is not None:
                state.ws, state.key = ws, key
        y = torch.empty(batch, n_out, device=xc.device, dtype=torch.float32)
        kh = _host_floats(knots_host)
        with torch.cuda.device(xc.device):
            _lib.call('rvk_kan_layer_forward', _p(xc), _p(sw), _p(lw), _p(lb), kh, len(knots_host), batch, n_in, n_out,
                      int(act), _p(y), _p(ws), int(need_bwd) | (2 if prepared else 0), _stream())
        if need_bwd:
            ctx.save_for_backward(xc, y, sw, lw)
            ctx.ws, ctx.knots_host, ctx.act = ws, tuple(knots_host), int(act)
        return y

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, gy):
        xc, y, sw, lw = ctx.saved_tensors
        batch, n_in = xc.shape
        n_out = lw.shape[0]
        g = _f32c(gy)
        need_x = ctx.needs_input_grad[0]
        need_w = any(ctx.needs_input_grad[1:4])
        dx = torch.empty_like(xc) if need_x else None
        dsw = torch.zeros_like(sw) if need_w else None
        dlw = torch.zeros_like(lw) if need_w else None
        dlb = torch.zeros(n_out, device=xc.device, dtype=torch.float32) if need_w else None
        with torch.cuda.device(xc.device):
            _lib.call('rvk_kan_layer_backward', _p(xc), _p(y), _p(g), _p(sw), _p(lw), _host_floats(ctx.knots_host),
                      len(ctx.knots_host), batch, n_in, n_out, ctx.act, _p(dx), _p(dsw), _p(dlw), _p(dlb), _p(ctx.ws),
                      _stream())
        return dx, dsw, dlw, dlb, None, None, None


class KanLayerState:
    """Per-layer cache of the prepared weight workspace (see KanLayerFn.forward)."""

    def __init__(self):
        self.ws = None
        self.key = None


# --------------------------------------------------------------------------------------------- heads
class LinearFn(torch.autograd.Function):
    """y = epilogue(x W^T + b) with fused ReLU / dropout / clamp (heads.py:17-22, 38-43, 91-102)."""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, x, w, b, relu, drop_p, clamp):
        require_cuda(x, 'Linear head forward')
        xc, wc, bc = _f32c(x), _f32c(w), _f32c(b)
        batch, n_in = xc.shape
        n_out = wc.shape[0]
        lo, hi = (float(clamp[0]), float(clamp[1])) if clamp is not None else (0.0, 0.0)
        seed, offset = _rng_keys() if drop_p > 0.0 else (0, 0)
        y = torch.empty(batch, n_out, device=xc.device, dtype=torch.float32)
        with torch.cuda.device(xc.device):
            _lib.call('rvk_linear_forward', _p(xc), _p(wc), _p(bc), batch, n_in, n_out, int(relu), float(drop_p), seed,
                      offset, lo, hi, _p(y), _stream())
        if any(ctx.needs_input_grad[:3]):
            ctx.save_for_backward(xc, wc, y)
            ctx.cfg = (int(relu), float(drop_p), lo, hi)
        return y

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, gy):
        xc, wc, y = ctx.saved_tensors
        relu, drop_p, lo, hi = ctx.cfg
        batch, n_in = xc.shape
        n_out = wc.shape[0]
        g = _f32c(gy)
        dx = torch.empty_like(xc) if ctx.needs_input_grad[0] else None
        dw = torch.zeros_like(wc) if ctx.needs_input_grad[1] else None
        db = torch.zeros(n_out, device=xc.device, dtype=torch.float32) if ctx.needs_input_grad[2] else None
        ws = torch.empty(batch, n_out, device=xc.device, dtype=torch.float32)
        with torch.cuda.device(xc.device):
            _lib.call('rvk_linear_backward', _p(xc), _p(wc), _p(y), _p(g), batch, n_in, n_out, relu, drop_p, lo, hi,
                      _p(dx), 0, _p(dw), _p(db), _p(ws), _stream())
        return dx, dw, db, None, None, None


# --------------------------------------------------------------------------------------------- loss
class JointLossFn(torch.autograd.Function):
    """Fused JointLoss.forward (training/losses.py:139-181): returns [cls, ord, unc, kan, total]."""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, cls_logits, ord_logits, mu, log_var, kan, class_t, sev_t, alpha, gamma, lambda_ord, mu_unc, nu_kan):
        require_cuda(cls_logits, 'JointLoss.forward')
        dev = cls_logits.device
        batch, ncls = cls_logits.shape
        need = any(ctx.needs_input_grad[:5])
        t = lambda v: None if v is None else _f32c(v)
        cl, ol, m, lv, kn = t(cls_logits), t(ord_logits), t(mu), t(log_var), t(kan)
        ct = class_t.to(device=dev, dtype=torch.int64).contiguous()
        st = sev_t.to(device=dev, dtype=torch.float32).contiguous()      # the reference casts severities with .float() (losses.py:92,112)
        al = None if alpha is None else alpha.to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty(5, device=dev, dtype=torch.float32)
        sums = torch.empty(4, device=dev, dtype=torch.float32)
        mk = lambda v: torch.empty_like(v) if (need and v is not None) else None
        d_cls, d_ord, d_mu, d_lv, d_kan = mk(cl), mk(ol), mk(m), mk(lv), mk(kn)
        with torch.cuda.device(dev):
            _lib.call('rvk_joint_loss_forward', _p(cl), ncls, _p(ol), _p(m), _p(lv), _p(kn), _p(ct), _p(st), _p(al),
                      float(gamma), float(lambda_ord), float(mu_unc), float(nu_kan), batch, _p(sums), _p(out),
                      _p(d_cls), _p(d_ord), _p(d_mu), _p(d_lv), _p(d_kan), _stream())
        if need:
            ctx.local = (d_cls, d_ord, d_mu, d_lv, d_kan)
            ctx.weights = (1.0, float(lambda_ord), float(mu_unc), float(mu_unc), float(nu_kan))
        return out

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, g5):
        g = _f32c(g5)
        terms = (0, 1, 2, 2, 3)
        grads = []
        with torch.cuda.device(g.device):
            for i, (loc, w, term) in enumerate(zip(ctx.local, ctx.weights, terms)):
                if loc is None or not ctx.needs_input_grad[i]:
                    grads.append(None)
                    continue
                dst = torch.empty_like(loc)
                _lib.call('rvk_joint_loss_backward', _p(loc), _p(g), term, w, _p(dst), loc.numel(), _stream())
                grads.append(dst)
        return (*grads, None, None, None, None, None, None, None)


# --------------------------------------------------------------------------------------------- encoder
class EncoderState:
    """Per-module cache for the trunk: bf16 weight buffer, inference workspace, pointer tables."""

    def __init__(self):
        self.wbuf = None
        self.wkey = None
        self.infer_ws = None
        self.infer_key = None
        # uint8 input: (pixel / 255 - mean) / std, ImageNet statistics as in the reference's transforms
        self.pixel_mean = (0.485, 0.456, 0.406)
        self.pixel_std = (0.229, 0.224, 0.225)

    def pixel_scale_shift(self):
        scale = [1.0 / (255.0 * s) for s in self.pixel_std]
        shift = [-m / s for m, s in zip(self.pixel_mean, self.pixel_std)]
        return scale, shift

    def param_table(self, tensors):
        return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    def weights(self, params, training: bool, device, key_params=None):
        # the key comes from the tensors the CALLER owns (`key_params`): for a non-fp32 module `params` are fresh
        # `.float()` copies that sit at version 0 (and usually at the same address) on every call
        kp = params if key_params is None else key_params
        key = (training, tuple(p.data_ptr() for p in kp), tuple(p._version for p in kp), tuple(p.dtype for p in kp))
        if self.wbuf is None or self.wkey != key:
            lib = _lib.load()
            nbytes = lib.rvk_encoder_weight_bytes(int(training))
            if self.wbuf is None or self.wbuf.numel() != nbytes or self.wbuf.device != device:
                self.wbuf = torch.empty(nbytes, device=device, dtype=torch.uint8)
            _lib.call('rvk_encoder_prepare_weights', self.param_table(params), _p(self.wbuf), int(training), _stream())
            self.wkey = key
        return self.wbuf

    def inference_workspace(self, batch: int, chunk: int, device):
        key = (batch, chunk, device)
        if self.infer_ws is None or self.infer_key != key:
            nbytes = _lib.load().rvk_encoder_workspace_bytes(batch, 0, chunk)
            self.infer_ws = torch.empty(nbytes, device=device, dtype=torch.uint8)
            self.infer_key = key
        return self.infer_ws


class EncoderFn(torch.autograd.Function):
    """DeiT-Tiny trunk forward/backward (replaces timm's VisionTransformer.forward that
    reference models/backbone.py:23-25 calls).  `params`: the 150 trunk tensors in timm order."""

    @staticmethod
    @custom_fwd(device_type='cuda')
    def forward(ctx, state, images, *params):
        require_cuda(images, 'DeiTTinyBackbone.forward')
        if images.dim() != 4 or tuple(images.shape[1:]) != (3, 224, 224):
            raise ValueError(f'expected images of shape (B, 3, 224, 224), got {tuple(images.shape)}')
        for p in params:
            require_cuda(p, 'DeiTTinyBackbone parameters')
        # bf16 images are consumed as they are (the trunk rounds pixels to bf16 first anyway); uint8 pixels are normalised
        # inside the patch gather with state.pixel_norm (ToTensor + Normalize folded in); anything else runs as fp32
        img_bf16 = images.dtype == torch.bfloat16
        img_u8 = images.dtype == torch.uint8
        img = images.contiguous() if (img_bf16 or img_u8) else _f32c(images)
        owner_params = params
        params = tuple(p.float() if p.dtype != torch.float32 else p for p in params)
        batch = img.shape[0]
        dev = img.device
        training = any(ctx.needs_input_grad[2:])      # callers pass detached tensors under torch.no_grad()
        chunk = CHUNK_IMAGES
        feats = torch.empty(batch, 192, device=dev, dtype=torch.float32)
        pc = [p.detach() for p in params]
        with torch.cuda.device(dev):
            wbuf = state.weights(pc, training, dev, key_params=owner_params)
            if training:
                nbytes = _lib.load().rvk_encoder_workspace_bytes(batch, 1, chunk)
                ws = torch.empty(nbytes, device=dev, dtype=torch.uint8)
            else:
                ws = state.inference_workspace(batch, chunk, dev)
            if img_u8:
                scale, shift = state.pixel_scale_shift()
                _lib.call('rvk_encoder_forward_u8', state.param_table(pc), _p(wbuf), _p(img), _host_floats(scale),
                          _host_floats(shift), batch, int(training), chunk, _p(ws), _p(feats), _stream())
            else:
                _lib.call('rvk_encoder_forward_bf16' if img_bf16 else 'rvk_encoder_forward', state.param_table(pc), _p(wbuf),
                          _p(img), batch, int(training), chunk, _p(ws), _p(feats), _stream())
        if training:
            ctx.state, ctx.ws, ctx.wbuf, ctx.batch, ctx.chunk = state, ws, wbuf, batch, chunk
            ctx.params = pc
        return feats

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, dfeat):
        g = _f32c(dfeat)
        params = ctx.params
        sizes = [p.numel() for p in params]
        flat = torch.zeros(sum(sizes), device=g.device, dtype=torch.float32)
        views, off = [], 0
        for p, n in zip(params, sizes):
            views.append(flat[off:off + n].view(p.shape))
            off += n
        from . import dist as rdist
        ov = rdist.overlap_state()
        with torch.cuda.device(g.device):
            if ov is None or ctx.batch > ctx.chunk:
                _lib.call('rvk_encoder_backward', ctx.state.param_table(params), _p(ctx.wbuf), _p(ctx.ws), _p(g), ctx.batch,
                          ctx.chunk, ctx.state.param_table(views), _stream())
            else:
                # data-parallel training: the gradient slice that is final after each stage range goes to NCCL while the next
                # range computes.  Parameter order = buffer order: [cls, pos, patch w/b | block 0 .. block 11 | norm w/b]
                ptab, gtab = ctx.state.param_table(params), ctx.state.param_table(views)
                block_start = lambda b: sum(sizes[:4 + 12 * b])
                hi = len(flat)
                for s_begin, s_end, first_final in rdist.bucket_stage_ranges(ov['buckets']):
                    _lib.call('rvk_encoder_backward_range', ptab, _p(ctx.wbuf), _p(ctx.ws), _p(g), ctx.batch, ctx.chunk, gtab,
                              s_begin, s_end, _stream())
                    lo = 0 if first_final < 0 else block_start(first_final)
                    rdist.launch_bucket(flat, lo, hi)
                    hi = lo
        ctx.ws = None
        return (None, None, *views)


def encoder_traced_forward(state: EncoderState, images: torch.Tensor, params):
    """Trunk forward in the activation-saving (training) layout WITHOUT autograd, returning the features and a function
    `saved(block, which)` that views the tensors the kernels left in the workspace (rvk_encoder_saved_offset).  This is the
    hook / explainability mode: the fused trunk never runs `blocks[i].attn` etc. as modules, so their forward
    hooks are fed from here."""
    require_cuda(images, 'DeiTTinyBackbone (hook mode)')
    if images.dim() != 4 or tuple(images.shape[1:]) != (3, 224, 224):
        raise ValueError(f'expected images of shape (B, 3, 224, 224), got {tuple(images.shape)}')
    lib = _lib.load()
    img = _f32c(images)
    batch, dev = img.shape[0], img.device
    pc = [(p.detach().float() if p.dtype != torch.float32 else p.detach()) for p in params]
    feats = torch.empty(batch, 192, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        wbuf = state.weights(pc, True, dev, key_params=list(params))
        ws = torch.empty(lib.rvk_encoder_workspace_bytes(batch, 1, batch), device=dev, dtype=torch.uint8)
        _lib.call('rvk_encoder_forward', state.param_table(pc), _p(wbuf), _p(img), batch, 1, batch, _p(ws), _p(feats), _stream())
    widths = {0: (192, torch.float32), 1: (192, torch.bfloat16), 2: (576, torch.bfloat16), 3: (192, torch.bfloat16),
              4: (192, torch.float32), 5: (192, torch.bfloat16), 6: (768, torch.bfloat16), 7: (768, torch.bfloat16)}
    rows = batch * 197

    def saved(block: int, which: int) -> torch.Tensor:
        off = lib.rvk_encoder_saved_offset(batch, block, which)
        if off < 0:
            raise ValueError(f'no saved tensor (block {block}, which {which})')
        width, dtype = widths[which]
        nbytes = rows * width * (4 if dtype == torch.float32 else 2)
        return ws[off:off + nbytes].view(dtype).view(batch, 197, width)

    return feats, saved, ws


def attention_probabilities(qkv: torch.Tensor) -> torch.Tensor:
    """softmax(q k^T / 8) per (image, head) from a saved qkv tensor (B, 197, 576) bf16 -> (B, 3, 197, 197) fp32."""
    require_cuda(qkv, 'attention_probabilities')
    q = qkv.contiguous()
    batch = q.shape[0]
    out = torch.empty(batch, 3, 197, 197, device=q.device, dtype=torch.float32)
    with torch.cuda.device(q.device):
        _lib.call('rvk_attention_probs', _p(q), _p(out), batch, _stream())
    return out


# --------------------------------------------------------------------------------------------- fused inference tail
class HeadsFusedState:
    """Cache of the repacked head / KAN weights for the one-kernel inference tail (rvk_heads_fused)."""

    def __init__(self):
        self.ws = None
        self.key = None


def heads_fused(state: HeadsFusedState, features: torch.Tensor, params, knots_host):
    """All four heads of RoViTKAN.forward (eval mode, reference rovit_kan.py:96-124) in one kernel launch.
    `params`: the 23 tensors in the order documented in include/rovitkan.h."""
    require_cuda(features, 'RoViTKAN heads')
    f = _f32c(features)
    batch, dev = f.shape[0], f.device
    pc = [_f32c(p.detach()) for p in params]
    key = (dev, tuple(p.data_ptr() for p in pc), tuple(p._version for p in params))
    lib = _lib.load()
    with torch.cuda.device(dev):
        if state.ws is None or state.key != key:
            if state.ws is None or state.ws.device != dev:
                state.ws = torch.empty(lib.rvk_heads_fused_workspace_floats(), device=dev, dtype=torch.float32)
            table = (C.c_void_p * len(pc))(*[p.data_ptr() for p in pc])
            _lib.call('rvk_heads_fused_prepare', table, _p(state.ws), _stream())
            state.key = key
        cls = torch.empty(batch, 4, device=dev, dtype=torch.float32)
        ordl = torch.empty(batch, 3, device=dev, dtype=torch.float32)
        mu = torch.empty(batch, 1, device=dev, dtype=torch.float32)
        lv = torch.empty(batch, 1, device=dev, dtype=torch.float32)
        kan = torch.empty(batch, 1, device=dev, dtype=torch.float32)
        _lib.call('rvk_heads_fused', _p(f), _p(state.ws), _host_floats(knots_host), batch, _p(cls), _p(ordl), _p(mu), _p(lv),
                  _p(kan), _stream())
    return cls, ordl, mu, lv, kan


class HeadsTrainFn(torch.autograd.Function):
    """The whole multi-task tail of the TRAINING step as one forward and one backward kernel (north_star (c); reference
    rovit_kan.py:96-124 with heads.py:17-22, 38-43, 91-102 and kan.py:138-149).  Inputs: features (B,192) and the 23 head / KAN
    parameters in rvk_heads_fused_prepare order; outputs (cls_logits, ordinal_logits, mu, log_var, kan_severity).  Outputs the
    caller does not use receive no gradient, and the parameters behind them get `None` (stage gating, rovit_kan.py:77)."""

    @staticmethod
    @custom_fwd(device_type='cuda', cast_inputs=torch.float32)
    def forward(ctx, features, knots_host, drop_p, *params):
        require_cuda(features, 'RoViTKAN heads (training)')
        f = _f32c(features)
        batch, dev = f.shape[0], f.device
        pc = [_f32c(p) for p in params]
        lib = _lib.load()
        nws = lib.rvk_heads_fused_workspace_floats()
        ws = torch.empty(nws, device=dev, dtype=torch.float32)
        seed, offset = _rng_keys() if drop_p > 0.0 else (0, 0)
        mk = lambda n: torch.empty(batch, n, device=dev, dtype=torch.float32)
        cls, ordl, mu, lv, kan = mk(4), mk(3), mk(1), mk(1), mk(1)
        h, a1, a2 = mk(384), mk(64), mk(16)
        kh = _host_floats(knots_host)
        with torch.cuda.device(dev):
            table = (C.c_void_p * len(pc))(*[p.data_ptr() for p in pc])
            _lib.call('rvk_heads_fused_prepare', table, _p(ws), _stream())
            _lib.call('rvk_heads_train_forward', _p(f), _p(ws), kh, batch, float(drop_p), seed, offset, _p(cls), _p(ordl), _p(mu),
                      _p(lv), _p(kan), _p(h), _p(a1), _p(a2), _stream())
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(f, ws, h, a1, a2, lv, kan)
        ctx.cfg = (tuple(knots_host), float(drop_p), [tuple(p.shape) for p in pc])
        return cls, ordl, mu, lv, kan

    @staticmethod
    @custom_bwd(device_type='cuda')
    def backward(ctx, g_cls, g_ord, g_mu, g_lv, g_kan):
        f, ws, h, a1, a2, lv, kan = ctx.saved_tensors
        knots_host, drop_p, shapes = ctx.cfg
        batch, dev = f.shape[0], f.device
        gs = [None if g is None else _f32c(g) for g in (g_cls, g_ord, g_mu, g_lv, g_kan)]
        need = ctx.needs_input_grad
        # parameter blocks: 0-3 classification, 4-7 ordinal, 8-13 uncertainty, 14-22 KAN
        live = [gs[0] is not None] * 4 + [gs[1] is not None] * 4 + [gs[2] is not None or gs[3] is not None] * 6 + [gs[4] is not None] * 9
        # the live gradients are views of ONE flat buffer in parameter order: a data-parallel caller reduces them in place
        # with a single collective (dist.all_reduce_gradients), like the trunk's flat gradient
        numels = [int(torch.Size(shp).numel()) if (live[i] and need[3 + i]) else 0 for i, shp in enumerate(shapes)]
        # (an empty batch launches nothing: its gradients are zeros, not whatever the allocator hands out)
        flat_g = (torch.empty if batch > 0 else torch.zeros)(sum(numels), device=dev, dtype=torch.float32)
        grads, off = [], 0
        for n, shp in zip(numels, shapes):
            grads.append(flat_g[off:off + n].view(shp) if n else None)
            off += n
        dfeat = torch.empty_like(f)
        dws = torch.empty_like(ws)
        with torch.cuda.device(dev):
            gtable = (C.c_void_p * len(grads))(*[_p(g) for g in grads])
            _lib.call('rvk_heads_train_backward', _p(f), _p(ws), _host_floats(knots_host), batch, drop_p, _p(h), _p(a1), _p(a2),
                      _p(lv), _p(kan), _p(gs[0]), _p(gs[1]), _p(gs[2]), _p(gs[3]), _p(gs[4]), _p(dfeat), _p(dws), gtable, _stream())
        return (dfeat if need[0] else None, None, None, *grads)


def predict_decode(cls_logits, ordinal_logits, log_var):
    """RoViTKAN.predict epilogue in one launch (rvk_predict_decode): (class, class_probs, ordinal_probs, ordinal_severity, std)."""
    require_cuda(cls_logits, 'RoViTKAN.predict')
    cl = _f32c(cls_logits)
    batch, ncls = cl.shape
    dev = cl.device
    ol = None if ordinal_logits is None else _f32c(ordinal_logits)
    lv = None if log_var is None else _f32c(log_var)
    idx = torch.empty(batch, device=dev, dtype=torch.int64)
    probs = torch.empty(batch, ncls, device=dev, dtype=torch.float32)
    oprobs = torch.empty(batch, ncls, device=dev, dtype=torch.float32) if ol is not None else None
    osev = torch.empty(batch, 1, device=dev, dtype=torch.float32) if ol is not None else None
    std = torch.empty(batch, 1, device=dev, dtype=torch.float32) if lv is not None else None
    with torch.cuda.device(dev):
        _lib.call('rvk_predict_decode', _p(cl), ncls, _p(ol), _p(lv), batch, _p(idx), _p(probs), _p(oprobs), _p(osev), _p(std), _stream())
    return idx, probs, oprobs, osev, std
