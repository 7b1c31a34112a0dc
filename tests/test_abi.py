"""The C-ABI library loads on a CPU-only box and exports every symbol include/branchmpc.h declares; the ctypes
mirror of the structs has the layout the C compiler gives them.  (No compute calls: those need a GPU.)"""
import ctypes as C
import os
import re
import subprocess

import pytest

from tests.helpers import ROOT
from _bmpc import abi

HEADER = os.path.join(ROOT, "include", "branchmpc.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bmpc_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_and_loads():
    import __graft_entry__ as entry
    entry.build()
    lib = abi.load_library()
    assert lib.bmpc_version() == 201


def test_every_declared_symbol_is_exported_and_bound():
    lib = abi.load_library()
    declared = _declared_functions()
    bound = sorted(name for name, _, _ in abi.SYMBOLS)
    assert declared == bound, "header and ctypes mirror disagree"
    for name in declared:
        assert getattr(lib, name) is not None


def test_struct_layout_matches_the_c_compiler(tmp_path):
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "branchmpc.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(bmpc_config), sizeof(bmpc_outputs),'
                   'offsetof(bmpc_config, dt), offsetof(bmpc_config, Q), offsetof(bmpc_config, row_f),'
                   'offsetof(bmpc_config, max_iter), offsetof(bmpc_config, batch_capacity));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(abi.Config), C.sizeof(abi.Outputs), abi.Config.dt.offset, abi.Config.Q.offset,
            abi.Config.row_f.offset, abi.Config.max_iter.offset, abi.Config.batch_capacity.offset]
    assert got == want


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(abi, "_lib", None)
    monkeypatch.setattr(abi, "library_path", lambda: str(tmp_path / "libbranchmpc.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        abi.load_library()


def test_error_codes_without_a_device():
    """Argument validation happens before any CUDA call, so it can be exercised on a CPU box."""
    from _bmpc import scenarios
    lib = abi.load_library()
    cfg = scenarios.highway_config()
    cfg.NB = 7
    h = C.c_void_p()
    assert lib.bmpc_create(C.byref(cfg), C.byref(h)) == abi.E_INVALID
    assert b"NB" in lib.bmpc_last_error(None)
    cfg = scenarios.highway_config()
    cfg.controller = abi.CTRL_ROBUST
    cfg.dR[0] = 1.0          # robustMPC with input-rate costs is not built
    assert lib.bmpc_create(C.byref(cfg), C.byref(h)) == abi.E_UNSUPPORTED
    cfg = scenarios.highway_config()
    cfg.controller = 17
    assert lib.bmpc_create(C.byref(cfg), C.byref(h)) == abi.E_INVALID
    assert lib.bmpc_create(None, C.byref(h)) == abi.E_INVALID
