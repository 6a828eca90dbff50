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
    assert lib.sd_abi_version() == 2
    assert isinstance(lib.sd_last_error(), bytes)
    assert lib.sd_set_impl(99) != 0
    assert b"sd_set_impl" in lib.sd_last_error()


def test_struct_layout_matches_c(tmp_path):
    """sizeof / offsetof of every struct that crosses the boundary, as gcc lays them out from include/sd_b200.h,
    against the ctypes mirrors in sd_b200/_native.py."""
    import subprocess
    structs = {"sd_conv_args": _native.ConvArgs, "sd_wgrad_args": _native.WgradArgs, "sd_pack_entry": _native.PackEntry,
               "sd_adam_entry": _native.AdamEntry}
    cname = {"inp": "in"}
    src = ["#include <stdio.h>", "#include <stddef.h>", '#include "sd_b200.h"', "int main(void) {"]
    for cn, st in structs.items():
        src.append('printf("%s %%zu\\n", sizeof(%s));' % (cn, cn))
        for f, _ in st._fields_:
            src.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cn, f, cn, cname.get(f, f)))
    src.append("return 0; }")
    c = tmp_path / "abi.c"
    c.write_text("\n".join(src))
    exe = str(tmp_path / "abi")
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", exe])
    got = dict(l.split() for l in subprocess.check_output([exe], text=True).splitlines())
    for cn, st in structs.items():
        assert int(got[cn]) == ctypes.sizeof(st), cn
        for f, _ in st._fields_:
            assert int(got["%s.%s" % (cn, f)]) == getattr(st, f).offset, (cn, f)
