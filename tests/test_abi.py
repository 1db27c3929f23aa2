"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/rovitkan.h declares, the ctypes binding covers the same set, and the product path refuses
to run without CUDA (no silent fallback).  No compute calls are made here."""

import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'rovitkan.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(rvk_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib_path():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    from rovitkan_b200 import _lib
    return _lib.LIB_PATH


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f'{n} declared in rovitkan.h but not exported by librovitkan.so'


def test_binding_covers_header(lib_path):
    from rovitkan_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.rvk_abi_version() == 1
    assert b'ok' == lib.rvk_strerror(0)
    assert lib.rvk_kan_layer_workspace_floats(192, 64, 0) == 2 * 192 * 8 * 64 + 64
    assert lib.rvk_encoder_workspace_bytes(0, 0, 0) == 0
    assert lib.rvk_encoder_weight_bytes(1) > lib.rvk_encoder_weight_bytes(0) > 5_400_000 * 2


def test_only_c_symbols_are_exported(lib_path):
    out = subprocess.run(['nm', '-D', '--defined-only', lib_path], capture_output=True, text=True).stdout
    exported = [l.split()[-1] for l in out.splitlines() if ' T ' in l]
    assert exported and all(s.startswith('rvk_') for s in exported), exported


def test_sass_contains_blackwell_tensor_and_tma_instructions(lib_path):
    out = subprocess.run(['cuobjdump', '-sass', lib_path], capture_output=True, text=True).stdout
    for mnemonic in ('UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM'):
        assert mnemonic in out, f'{mnemonic} missing from SASS: not a tcgen05/TMA build'


def test_fused_mlp_variants_are_validated_before_any_cuda_call(lib_path):
    """cta_group of rvk_mlp_fused / rvk_attn_proj_mlp_fused: 1 = single CTAs, 2 = CTA pairs, 4 = CTA pairs with two row tiles in flight
    (needs the folded projection, csrc/mlp_fused2.cuh); anything else, or 4 without the projection, is RVK_ERR_BAD_ARG.  The check
    sits in front of the launch, so it runs without a GPU (the pointers are never dereferenced)."""
    from rovitkan_b200 import _lib
    fake = 4096
    with pytest.raises(_lib.RovitKanError, match='bad argument'):
        _lib.call('rvk_mlp_fused', *([fake] * 10), 1e-6, fake, 128, 4, 0)
    with pytest.raises(_lib.RovitKanError, match='bad argument'):
        _lib.call('rvk_attn_proj_mlp_fused', *([fake] * 13), 1e-6, fake, 128, 3, 0)
    _lib.call('rvk_attn_proj_mlp_fused', *([fake] * 13), 1e-6, fake, 0, 4, 0)      # an empty batch is a no-op, not an error


def test_module_mirror_matches_reference_layout_and_rejects_cpu():
    from rovitkan_b200.models import KANLayer, RoViTKAN
    m = RoViTKAN(pretrained=False)
    want = [l.split(' ', 1) for l in open(os.path.join(ROOT, 'tests', 'golden', 'state_dict_keys.txt')).read().splitlines()]
    sd = m.state_dict()
    assert list(sd) == [k for k, _ in want]
    assert all(str(tuple(sd[k].shape)) == s for k, s in want)
    assert m.count_parameters()['total'] == 5706394
    assert RoViTKAN(embed_dim=192, pretrained=False).classification_head.fc1.in_features == 192   # scripts/train.py:88
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        m(torch.randn(1, 3, 224, 224))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        KANLayer(8, 4)(torch.randn(2, 8))
    with pytest.raises(NotImplementedError):
        KANLayer(8, 4, num_knots=7)


def test_reference_config_object_is_accepted():
    class _NS:
        pass
    cfg = _NS(); cfg.model = _NS(); cfg.data = _NS()
    cfg.model.embed_dim, cfg.model.hidden_dim, cfg.model.kan_layers = 192, 128, [192, 64, 16, 1]
    cfg.model.kan_num_knots, cfg.model.kan_degree, cfg.model.dropout, cfg.model.pretrained = 5, 3, 0.3, False
    cfg.data.num_classes = 4
    from rovitkan_b200.models import RoViTKAN
    m = RoViTKAN(cfg)
    assert m.kan_module.layers_dims == [192, 64, 16, 1] and m.curriculum_stage == 4


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'rovit-kan-interpretable-vision-transformer-for-rose-disease-severity-estimation_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), os.path.join(dirpath, f)


def test_pretrained_true_needs_weights_or_an_explicit_opt_in(tmp_path, monkeypatch):
    """pretrained=True (the reference default) must not silently train from scratch."""
    from rovitkan_b200.models import deit
    from rovitkan_b200.models import RoViTKAN
    monkeypatch.delenv(deit.PRETRAINED_ENV, raising=False)
    try:
        import timm  # noqa: F401
        have_timm = True
    except ImportError:
        have_timm = False
    if not have_timm:
        with pytest.raises(RuntimeError, match='ROVITKAN_PRETRAINED'):
            RoViTKAN(pretrained=True)
    monkeypatch.setenv(deit.PRETRAINED_ENV, 'random')
    with pytest.warns(UserWarning, match='randomly initialised'):
        RoViTKAN(pretrained=True)
    # a timm-style checkpoint (with a classifier head that num_classes=0 drops) is loaded strictly
    src = deit.create_model(pretrained=False)
    sd = {k: torch.full_like(v, 0.25) for k, v in src.state_dict().items()}
    sd['head.weight'] = torch.zeros(1000, 192)
    sd['head.bias'] = torch.zeros(1000)
    ck = tmp_path / 'deit_tiny.pth'
    torch.save({'model': sd}, ck)
    monkeypatch.setenv(deit.PRETRAINED_ENV, str(ck))
    m = RoViTKAN(pretrained=True)
    assert all(bool((p == 0.25).all()) for p in m.backbone.model.parameters())


def test_trunk_weight_cache_key_follows_the_owner_parameters():
    """for a non-fp32 module the trunk received fresh .float() copies (version 0, recycled address) and
    kept stale bf16 weights after optimizer.step(); the key is now taken from the owner's parameters."""
    from rovitkan_b200 import ops
    st = ops.EncoderState()
    seen = []
    st.wbuf = torch.empty(1)

    class _Lib:
        @staticmethod
        def rvk_encoder_weight_bytes(training):
            return 1
    real_load, real_call, real_stream = ops._lib.load, ops._lib.call, ops._stream
    ops._lib.load = lambda: _Lib
    ops._lib.call = lambda name, *a: seen.append(name)
    ops._stream = lambda: 0
    try:
        owner = [torch.zeros(4, dtype=torch.bfloat16)]
        for _ in range(2):
            st.weights([owner[0].float()], False, torch.device('cpu'), key_params=owner)
        assert seen == ['rvk_encoder_prepare_weights']            # second call: cache hit
        owner[0].add_(1)                                          # what optimizer.step() does
        st.weights([owner[0].float()], False, torch.device('cpu'), key_params=owner)
        assert len(seen) == 2
    finally:
        ops._lib.load, ops._lib.call, ops._stream = real_load, real_call, real_stream


HOUSEKEEPING = {'rvk_abi_version', 'rvk_strerror', 'rvk_last_error', 'rvk_device_check', 'rvk_launch_count', 'rvk_stream_check',
                'rvk_set_side_stream', 'rvk_debug_mbar_timeout', 'rvk_gemm_timing_enable', 'rvk_gemm_timing_collect',
                'rvk_gemm_timing_kind', 'rvk_timing_enable', 'rvk_timing_collect', 'rvk_timing_kind', 'rvk_timing_kind_name',
                'rvk_debug_set_mlp_trace', 'rvk_debug_set_attn_trace'}


def test_every_compute_entry_point_refuses_null_pointers_with_a_status(lib_path):
    """Error behaviour of the boundary: the C entry points never throw and never dereference an argument
    they have not checked -- null pointers come back as RVK_ERR_BAD_ARG, which the binding turns into RovitKanError (a
    RuntimeError, as the reference's callers would see from torch).  The checks sit in front of every CUDA call, so this
    runs without a GPU."""
    from rovitkan_b200 import _lib
    lib = _lib.load()
    assert issubclass(_lib.RovitKanError, RuntimeError)
    checked = 0
    for name, (res, args) in _lib.SIGNATURES.items():
        if name in HOUSEKEEPING or res is not ctypes.c_int:
            continue
        vals = [None if a is ctypes.c_void_p else (0.5 if a in (ctypes.c_float, ctypes.c_double) else 1) for a in args]
        assert getattr(lib, name)(*vals) == 1, f'{name}: null pointers must give RVK_ERR_BAD_ARG'
        with pytest.raises(_lib.RovitKanError, match='bad argument'):
            _lib.call(name, *vals)
        checked += 1
    assert checked >= 30
    # size queries: a negative or null size table is refused the same way (no allocation request derived from it)
    assert lib.rvk_optimizer_state_floats(1, None) == -1


def test_knot_vector_other_than_the_reference_linspace_is_refused_at_the_boundary(lib_path):
    """Every KAN kernel evaluates the closed uniform-knot cubic segments of the reference's linspace(-1, 1, 11) buffer
    (models/kan.py:59-60); another knot vector, which the reference's Cox-de Boor recursion would accept, is refused with a
    message instead of being evaluated as a different spline.  The check precedes the launch: no pointer is dereferenced."""
    from rovitkan_b200 import _lib
    fake = 4096
    good = (ctypes.c_float * 11)(*[-1.0 + 0.2 * i for i in range(11)])
    bad = (ctypes.c_float * 11)(*[-1.0 + 0.2 * i + (0.01 if i == 4 else 0.0) for i in range(11)])
    with pytest.raises(_lib.RovitKanError, match='linspace'):
        _lib.call('rvk_kan_layer_forward', fake, fake, fake, fake, bad, 11, 8, 192, 64, 0, fake, fake, 0, 0)
    with pytest.raises(_lib.RovitKanError, match='not supported'):
        _lib.call('rvk_kan_layer_forward', fake, fake, fake, fake, good, 9, 8, 192, 64, 0, fake, fake, 0, 0)
    _lib.call('rvk_kan_layer_forward', fake, fake, fake, fake, good, 11, 0, 192, 64, 0, fake, fake, 0, 0)   # empty batch: no-op
