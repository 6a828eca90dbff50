"""CPU: the C-ABI library builds/loads and exports every symbol include/sd_b200.h declares;
the ctypes binding covers all of them (no compute calls here)."""
import ctypes
import os
import re

from sd_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert "sd_conv_fwd" in syms and "sd_clip_dz" in syms and len(syms) >= 25


def test_library_exports_every_declared_symbol():
    assert os.path.isfile(_native.LIB_PATH), "build with `make -C speech-decoding_b200/csrc`"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), "missing export " + s


def test_ctypes_binding_covers_header():
    bound = set(_native.SIGNATURES) | {"sd_last_error"}
    assert set(declared_symbols()) == bound


def test_abi_version_and_error_string():
    lib = _native.lib()
    assert lib.sd_abi_version() == 1
    assert isinstance(lib.sd_last_error(), bytes)
    assert lib.sd_set_impl(99) != 0
    assert b"sd_set_impl" in lib.sd_last_error()


def test_struct_layout_matches_c():
    # sizeof() as the C compiler lays the structs out (9 pointers + 12 ints; 6 pointers + 9 ints + 4 i64 + int)
    assert ctypes.sizeof(_native.ConvArgs) == 9 * 8 + 12 * 4 + 8      # + the trailing `affine` pointer
    assert ctypes.sizeof(_native.WgradArgs) == 6 * 8 + 9 * 4 + 4 + 4 * 8 + 8 + 16
    assert ctypes.sizeof(_native.PackEntry) == 3 * 8 + 6 * 4
