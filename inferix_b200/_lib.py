"""ctypes binding of ``libinferix_b200.so`` (the C ABI declared in ``include/inferix_b200.h``).

This is the only place Python touches the native library.  There is no fallback: if the shared object is
missing the import of any op raises ``NativeLibraryError`` telling the user to build it
(``python -c 'import __graft_entry__ as g; g.build()'`` or ``make -C inferix_b200/csrc``).
Errors coming back over the ABI are mapped to the exception types the reference raises at the same spots
(SURVEY §8b): ValueError / IndexError / KeyError / RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libinferix_b200.so"

IFX_KV_MAX_PLAN_PAGES = 32

IFX_OK, IFX_ERR_INVALID, IFX_ERR_BOUNDS, IFX_ERR_HANDLE, IFX_ERR_OOM, IFX_ERR_CUDA, IFX_ERR_UNSUPPORTED = range(7)
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_GATE_RES, EPI_BIAS_GELU_ERF, EPI_BIAS_F32 = 0, 1, 2, 3, 4


class NativeLibraryError(RuntimeError):
    pass


class KvPlan(C.Structure):
    _fields_ = [
        ("local_start", C.c_int64),
        ("local_end", C.c_int64),
        ("global_end", C.c_int64),
        ("num_evicted", C.c_int64),
        ("num_pages", C.c_int32),
        ("pages", C.c_int32 * IFX_KV_MAX_PLAN_PAGES),
        ("first_offset", C.c_int32),
    ]


class RopeGrid(C.Structure):
    _fields_ = [
        ("frames", C.c_int32),
        ("height", C.c_int32),
        ("width", C.c_int32),
        ("start_frame", C.c_int32),
        ("hw_offset", C.c_int32),
        ("hw_count", C.c_int32),
    ]


IFX_MAX_PEERS = 8
IFX_PEER_HANDLE_BYTES = 64
IFX_ATTN_MAX_EXTENTS = 32
IFX_SP_STORE, IFX_SP_OVERLAP = 0, 1
IFX_Q8_E4M3, IFX_Q8_INT8 = 0, 1


class PeerDst(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32),
        ("k", C.c_void_p * IFX_MAX_PEERS), ("v", C.c_void_p * IFX_MAX_PEERS),
        ("flags", C.c_void_p * IFX_MAX_PEERS),
        ("epoch", C.c_int64),
        ("local_only", C.c_int32),
    ]


class WanBlockWeights(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("ffn_dim", C.c_int32), ("heads", C.c_int32), ("head_dim", C.c_int32),
        ("eps", C.c_float),
        ("qkv_w", C.c_void_p), ("qkv_b", C.c_void_p),
        ("norm_q_w", C.c_void_p), ("norm_k_w", C.c_void_p),
        ("o_w", C.c_void_p), ("o_b", C.c_void_p),
        ("norm3_w", C.c_void_p), ("norm3_b", C.c_void_p),
        ("cq_w", C.c_void_p), ("cq_b", C.c_void_p), ("cnorm_q_w", C.c_void_p),
        ("co_w", C.c_void_p), ("co_b", C.c_void_p),
        ("ffn1_w", C.c_void_p), ("ffn1_b", C.c_void_p),
        ("ffn2_w", C.c_void_p), ("ffn2_b", C.c_void_p),
    ]


class WanBlockIO(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("rows", C.c_int64), ("tokens_per_frame", C.c_int64),
        ("mod", C.c_void_p), ("freqs", C.c_void_p), ("grid", RopeGrid),
        ("kv", C.c_void_p), ("current_start", C.c_int64), ("sink_tokens", C.c_int64), ("windowed", C.c_int32),
        ("cross_k", C.c_void_p), ("cross_v", C.c_void_p), ("text_len", C.c_int64),
        ("ws_h", C.c_void_p), ("ws_qkv", C.c_void_p), ("ws_q", C.c_void_p), ("ws_attn", C.c_void_p),
        ("ws_ffn", C.c_void_p),
    ]


_vp, _i64, _i32, _f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float

# name -> (restype, argtypes); mirrors include/inferix_b200.h one to one (tests check the export list).
SIGNATURES = {
    "ifx_last_error": (C.c_char_p, []),
    "ifx_abi_version": (C.c_int, []),
    "ifx_launch_count": (C.c_uint64, []),
    "ifx_reset_launch_count": (None, []),
    "ifx_prof_enable": (None, [_i32]),
    "ifx_prof_reset": (None, []),
    "ifx_prof_read": (C.c_int, [C.c_char_p, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "ifx_prof_labels": (C.c_int, [C.c_char_p, _i32]),
    "ifx_kv_create": (C.c_int, [C.POINTER(_vp), _vp, _vp, _i32, _i32, _i32, _i32]),
    "ifx_kv_rebind": (C.c_int, [_vp, _vp, _vp]),
    "ifx_kv_destroy": (C.c_int, [_vp]),
    "ifx_kv_reset": (C.c_int, [_vp]),
    "ifx_kv_plan_append": (C.c_int, [_vp, _i64, _i64, _i64, _i32, C.POINTER(KvPlan)]),
    "ifx_kv_state": (C.c_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i32), C.POINTER(_i32), _i32]),
    "ifx_kv_map": (C.c_int, [_vp, _i64, C.POINTER(_vp), C.POINTER(_vp)]),
    "ifx_kv_export": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "ifx_kv_import": (C.c_int, [_vp, _vp, _vp, _i64, _i64, _vp]),
    "ifx_ln_modulate": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i64, _f32, _vp]),
    "ifx_quantize_fp8": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _f32, _vp]),
    "ifx_quantize_fp8_cols": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _vp, _vp]),
    "ifx_ln_modulate_fp8": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i64, _f32, _f32, _vp]),
    "ifx_gemm_fp8": (C.c_int, [_vp, _i64, _vp, _i64, _f32, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _vp,
                               _i64, _i64, _vp]),
    "ifx_quantize_rows": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _vp]),
    "ifx_ln_modulate_quant": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i64, _i32, _i64, _f32, _i32, _vp]),
    "ifx_gemm_q8": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i32, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _vp,
                              _i64, _i64, _vp]),
    "ifx_gemm_bf16": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i64, _i32, _i32, _i32, _vp, _i64, _vp, _i64,
                                _i64, _vp]),
    "ifx_qk_norm_rope_append": (C.c_int, [_vp, _i64, _vp, _vp, _vp, C.POINTER(RopeGrid), _vp, _i64, _vp,
                                          C.POINTER(KvPlan), _vp, _vp, _i64, _i32, _i32, _f32, _vp]),
    "ifx_peer_export": (C.c_int, [_vp, _vp, C.POINTER(_i64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "ifx_peer_open": (C.c_int, [_vp, C.POINTER(_vp)]),
    "ifx_peer_close": (C.c_int, [_vp]),
    "ifx_qk_norm_rope_append_peers": (C.c_int, [_vp, _i64, _vp, _vp, _vp, C.POINTER(RopeGrid), _vp, _i64, _vp,
                                                C.POINTER(KvPlan), C.POINTER(PeerDst), _i64, _i32, _i32, _f32, _vp]),
    "ifx_peer_wait": (C.c_int, [_vp, _i32, _i64, _i32, _vp]),
    "ifx_peer_push": (C.c_int, [_vp, C.POINTER(KvPlan), C.POINTER(PeerDst), _i32, _i32, _i32, _vp]),
    "ifx_kv_append": (C.c_int, [_vp, C.POINTER(KvPlan), _vp, _vp, _i64, _i64, _vp]),
    "ifx_kv_append_sp": (C.c_int, [_vp, C.POINTER(KvPlan), _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "ifx_rmsnorm": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i64, _i32, _f32, _vp]),
    "ifx_attention": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _f32, _vp]),
    "ifx_attention_gqa": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _vp, _i64, _i64, _i64, _i32, _i32, _i32, _f32, _vp]),
    "ifx_attention_partial": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i64, C.POINTER(_i64), _i32, _i64, _i32, _i32, _i32,
                                        _f32, _vp, _i64, _i32, _i32, _i32, _vp]),
    "ifx_attention_combine": (C.c_int, [_vp, _i32, _vp, _i64, _i64, _i32, _i32, _vp]),
    "ifx_attention_kv": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i64, _f32, _vp]),
    "ifx_attention_lse": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _f32, _vp]),
    "ifx_attention_extents": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i64, C.POINTER(_i64), _i32, _vp, _i64, _i64, _i32,
                                        _i32, _i32, _f32, _vp]),
    "ifx_attention_kv_wait": (C.c_int, [_vp, _i64, _vp, C.POINTER(KvPlan), _vp, _i32, _i64, _i32, _vp, _i64, _i64, _f32,
                                        _vp]),
    "ifx_magi_qkv_post": (C.c_int, [_vp, _i64, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32,
                                    _vp, _i64, _i32, _i64, _vp, _vp, _i64, _i32, _i64, _vp, _i64, _vp]),
    "ifx_head_layernorm": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _i32, _vp, _vp, _f32, _vp]),
    "ifx_gate_norm_residual": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _i64, _i64,
                                         _i32, _f32, _vp]),
    "ifx_silu_mul": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _vp]),
    "ifx_patchify": (C.c_int, [_vp, _i64, _i64, _i64, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "ifx_sinusoidal_embedding": (C.c_int, [_vp, _i32, _i32, _vp, _vp]),
    "ifx_linear_small": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _i32, _i64, _i64, _vp]),
    "ifx_unpatchify_x0": (C.c_int, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32,
                                    _i32, _vp, _vp, _vp]),
    "ifx_add_noise": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i64, _vp, _vp]),
    "ifx_wan_block_forward": (C.c_int, [C.POINTER(WanBlockWeights), C.POINTER(WanBlockIO), C.POINTER(KvPlan), _vp]),
    "ifx_wan_block_forward_sp": (C.c_int, [C.POINTER(WanBlockWeights), C.POINTER(WanBlockIO), C.POINTER(PeerDst), _i32,
                                           _i32, _i32, C.POINTER(KvPlan), _vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("INFERIX_B200_LIB", LIB_PATH))
    if not path.exists():
        raise NativeLibraryError(
            f"{path} not found: the CUDA extension is not built. Run `make -C inferix_b200/csrc` "
            "(or __graft_entry__.build()). inferix_b200 has no CPU / PyTorch fallback."
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    if lib.ifx_abi_version() != 1:
        raise NativeLibraryError(f"ABI version mismatch: library reports {lib.ifx_abi_version()}, binding expects 1")
    _lib = lib
    return lib


_EXC = {
    IFX_ERR_INVALID: ValueError,
    IFX_ERR_BOUNDS: IndexError,
    IFX_ERR_HANDLE: KeyError,
    IFX_ERR_OOM: MemoryError,
    IFX_ERR_CUDA: RuntimeError,
    IFX_ERR_UNSUPPORTED: NotImplementedError,
}


def check(status: int) -> None:
    if status != IFX_OK:
        msg = load().ifx_last_error().decode("utf-8", "replace")
        raise _EXC.get(status, RuntimeError)(msg)


def launch_count() -> int:
    return int(load().ifx_launch_count())


def reset_launch_count() -> None:
    load().ifx_reset_launch_count()


def prof_enable(on: bool) -> None:
    load().ifx_prof_enable(int(on))


def prof_reset() -> None:
    load().ifx_prof_reset()


def prof_read(prefix: str = ""):
    """(total_ms, launches) of all profiled kernels whose label starts with `prefix`."""
    ms, n = C.c_double(), C.c_uint64()
    check(load().ifx_prof_read(prefix.encode(), C.byref(ms), C.byref(n)))
    return ms.value, n.value


def prof_labels():
    buf = C.create_string_buffer(1 << 16)
    check(load().ifx_prof_labels(buf, len(buf)))
    return [s for s in buf.value.decode().split("\n") if s]
