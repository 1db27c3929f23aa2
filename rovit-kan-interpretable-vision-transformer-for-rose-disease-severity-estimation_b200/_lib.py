"""ctypes binding of librovitkan.so (the C ABI declared in include/rovitkan.h).

There is no fallback: if the library is missing or a call fails this raises.  Pointers travel as
integers (`tensor.data_ptr()`), the stream as `torch.cuda.current_stream().cuda_stream`.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'librovitkan.so')

_P, _I, _L, _F, _U64, _D = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64, C.c_double

# name -> (restype, argtypes); must list every function include/rovitkan.h declares
SIGNATURES = {
    'rvk_abi_version': (_I, []),
    'rvk_strerror': (C.c_char_p, [_I]),
    'rvk_last_error': (C.c_char_p, []),
    'rvk_device_check': (_I, []),
    'rvk_launch_count': (_L, []),
    'rvk_stream_check': (_I, [_P]),
    'rvk_set_side_stream': (None, [_I]),
    'rvk_debug_mbar_timeout': (_I, [_P]),
    'rvk_gemm_timing_enable': (None, [_I]),
    'rvk_gemm_timing_collect': (_I, [_P, _P]),
    'rvk_gemm_timing_kind': (_I, [_I, _P, _P]),
    'rvk_timing_enable': (None, [_I]),
    'rvk_timing_collect': (_I, []),
    'rvk_timing_kind': (_I, [_I, _P, _P, _P]),
    'rvk_timing_kind_name': (C.c_char_p, [_I]),
    'rvk_kan_layer_workspace_floats': (_L, [_I, _I, _I]),
    'rvk_kan_basis': (_I, [_P, _P, _I, _L, _P, _P]),
    'rvk_kan_layer_forward': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    'rvk_kan_layer_backward': (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    'rvk_linear_forward': (_I, [_P, _P, _P, _I, _I, _I, _I, _F, _U64, _U64, _F, _F, _P, _P]),
    'rvk_linear_backward': (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _F, _P, _I, _P, _P, _P, _P]),
    'rvk_heads_fused_workspace_floats': (_L, []),
    'rvk_heads_fused_prepare': (_I, [_P, _P, _P]),
    'rvk_heads_fused': (_I, [_P, _P, _P, _I, _P, _P, _P, _P, _P, _P]),
    'rvk_predict_decode': (_I, [_P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P]),
    'rvk_heads_train_forward': (_I, [_P, _P, _P, _I, _F, _U64, _U64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'rvk_heads_train_backward': (_I, [_P, _P, _P, _I, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    'rvk_joint_loss_forward': (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _P, _F, _F, _F, _F, _I, _P, _P, _P, _P, _P, _P, _P, _P]),
    'rvk_joint_loss_backward': (_I, [_P, _P, _I, _F, _P, _I, _P]),
    'rvk_optimizer_state_floats': (_L, [_I, _P]),
    'rvk_optimizer_step': (_I, [_I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _D, _D, _D, _D, _F, _F, _P, _P, _P]),
    'rvk_encoder_weight_bytes': (_L, [_I]),
    'rvk_encoder_workspace_bytes': (_L, [_I, _I, _I]),
    'rvk_encoder_prepare_weights': (_I, [_P, _P, _I, _P]),
    'rvk_encoder_forward': (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    'rvk_encoder_forward_bf16': (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P]),
    'rvk_encoder_forward_u8': (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    'rvk_encoder_saved_offset': (_L, [_I, _I, _I]),
    'rvk_attention_probs': (_I, [_P, _P, _I, _P]),
    'rvk_encoder_backward_range': (_I, [_P, _P, _P, _P, _I, _I, _P, _I, _I, _P]),
    'rvk_encoder_backward': (_I, [_P, _P, _P, _P, _I, _I, _P, _P]),
    'rvk_gemm_nt': (_I, [_I, _P, _L, _P, _L, _P, _L, _P, _L, _P, _L, _I, _I, _I, _P, _P, _P, _P, _I, _F, _P, _P, _P]),
    'rvk_mlp_fused': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _I, _I, _P]),
    'rvk_attn_proj_mlp_fused': (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _F, _P, _I, _I, _P]),
    'rvk_debug_set_mlp_trace': (None, [_P]),
    'rvk_debug_set_attn_trace': (None, [_P]),
    'rvk_gemm_tn': (_I, [_P, _L, _P, _L, _P, _L, _I, _I, _I, _F, _P, _P]),
    'rvk_attention_forward': (_I, [_P, _P, _P, _I, _P]),
    'rvk_attention_backward': (_I, [_P, _P, _P, _P, _P, _I, _P]),
    'rvk_layernorm_forward': (_I, [_P, _L, _P, _P, _F, _P, _I, _L, _P, _P, _I, _P]),
    'rvk_layernorm_backward': (_I, [_P, _I, _L, _P, _L, _P, _P, _P, _P, _P, _L, _P, _P, _P, _P, _I, _P]),
    'rvk_im2col': (_I, [_P, _P, _I, _P]),
    'rvk_cast_bf16': (_I, [_P, _P, _L, _P]),
}

_lib = None


class RovitKanError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load the CUDA library, building nothing: a missing .so is an error, not a fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RovitKanError(
            f'{LIB_PATH} not found. Build it with `python __graft_entry__.py` (or '
            f'`python {os.path.join(_HERE, "build.py")}`); there is no CPU or PyTorch fallback for this path.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    if lib.rvk_abi_version() != 1:
        raise RovitKanError(f'ABI version mismatch: library reports {lib.rvk_abi_version()}, binding expects 1')
    _lib = lib
    return lib


def check(status: int, what: str = '') -> None:
    if status != 0:
        lib = load()
        msg = lib.rvk_strerror(status).decode()
        detail = lib.rvk_last_error().decode()
        raise RovitKanError(f'{what or "rovitkan call"} failed: [{status}] {msg}' + (f' ({detail})' if detail else ''))


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
