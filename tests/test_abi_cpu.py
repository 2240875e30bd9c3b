"""CPU-side checks of the C-ABI boundary: the library loads without a GPU, exports exactly what the header declares,
the ctypes binding covers every export, and argument errors map to the reference's exception types."""
import ctypes
import re
from pathlib import Path

import pytest

from inferix_b200 import _lib

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "inferix_b200.h").read_text()


def declared_functions():
    # every `ifx_status name(` / `const char* name(` / `int name(` / `uint64_t name(` / `void name(` prototype
    names = re.findall(r"^(?:ifx_status|const char\*|int|uint64_t|void)\s+(ifx_\w+)\s*\(", HEADER, flags=re.M)
    assert len(names) >= 18
    return names


def test_library_loads_and_exports_header_symbols():
    lib = _lib.load()
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/inferix_b200.h but not exported"
    assert lib.ifx_abi_version() == 1


def test_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == sorted(declared_functions())


def test_no_torch_types_in_signatures():
    assert "torch" not in HEADER and "at::" not in HEADER and "c10" not in HEADER


def test_struct_layout_matches_header():
    # ifx_kv_plan: 4 x int64 + int32 + 32 x int32 + int32 (+ padding to 8)
    assert ctypes.sizeof(_lib.KvPlan) == 4 * 8 + 4 + 32 * 4 + 4
    assert ctypes.sizeof(_lib.RopeGrid) == 24
    # ifx_peer_dst: 2 x int32, 3 x 8 pointers, int64 epoch, int32 local_only (+ 4 padding)
    assert ctypes.sizeof(_lib.PeerDst) == 8 + 3 * 8 * 8 + 8 + 8
    assert _lib.PeerDst.epoch.offset == 200 and _lib.PeerDst.local_only.offset == 208


def test_errors_map_to_reference_exception_types():
    lib = _lib.load()
    with pytest.raises(KeyError):          # unknown handle -> KeyError like kvcache_manager.py free/get
        _lib.check(lib.ifx_kv_reset(None))
    h = ctypes.c_void_p()
    with pytest.raises(ValueError):        # bad geometry -> ValueError
        _lib.check(lib.ifx_kv_create(ctypes.byref(h), 0x1000, 0x2000, 0, 4, 1, 8))
    with pytest.raises(ValueError):
        _lib.check(lib.ifx_gemm_bf16(None, 8, None, 8, None, None, 8, 1, 8, 8, 0, None, 0, None, 0, 0, None))
    assert "null" in lib.ifx_last_error().decode()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setenv("INFERIX_B200_LIB", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.NativeLibraryError):
        _lib.load()
