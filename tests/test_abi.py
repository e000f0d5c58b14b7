"""CPU checks of the drop-in boundary: include/ne_b200.h <-> narvalengine_b200/abi.py <-> libnarval_b200.so.
No compute entry point is called (there is no GPU here); what IS checked is that the library loads, exports every
symbol the header declares, that the ctypes structs have the C layout, and that compute entry points fail loudly
(NE_B200_ERR_CUDA) instead of falling back to a CPU path when no device is visible."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from narvalengine_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ne_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ne_b200_[a-z0-9_]+)\s*\(", src)))


def test_header_and_ctypes_declare_the_same_symbols():
    assert header_symbols() == sorted(abi.SYMBOLS)


def test_library_loads_and_exports_every_declared_symbol():
    lib = abi.load_library()
    for name in header_symbols():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (ne_b200_[a-z0-9_]+)", out))
    assert exported == set(header_symbols())  # nothing undeclared leaks out either


def test_library_does_not_link_or_reference_the_oracle():
    out = subprocess.run(["ldd", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "narval_ref" not in out and "oracle" not in out
    syms = subprocess.run(["nm", "-D", abi.LIB_PATH], capture_output=True, text=True).stdout
    assert "neref_" not in syms


def test_struct_layouts_match_the_c_header(tmp_path):
    """Compile a probe against the header with gcc and compare sizeof/offsetof with the ctypes mirrors."""
    structs = {"ne_b200_texture": abi.Texture, "ne_b200_volume": abi.Volume, "ne_b200_material": abi.Material,
               "ne_b200_primitive": abi.Primitive, "ne_b200_scene_desc": abi.SceneDesc, "ne_b200_camera": abi.Camera,
               "ne_b200_hit": abi.Hit, "ne_b200_counters": abi.Counters}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ne_b200.h"', 'int main(void){']
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f, _ in ct._fields_:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, ct in structs.items():
        assert int(got[cname]) == C.sizeof(ct), cname
        for f, _ in ct._fields_:
            assert int(got[f"{cname}.{f}"]) == getattr(ct, f).offset, f"{cname}.{f}"


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "c.c"
    src.write_text('#include "ne_b200.h"\nint main(void){return NE_B200_API_VERSION == 1 ? 0 : 1;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "c.o")])


def test_no_cpu_fallback_without_a_device():
    lib = abi.load_library()
    if lib.ne_b200_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    h = C.c_void_p()
    rc = lib.ne_b200_create(0, C.byref(h))
    assert rc == abi.ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.ne_b200_last_error()
    # every compute entry point rejects a null context instead of computing anything
    assert lib.ne_b200_render(None, 8, 8, 0, 1, 1, 1, 0) == abi.ERR_INVALID
    assert lib.ne_b200_render_frame(None, None, 8, 8, 1, 1, 1, 0, None, None) == abi.ERR_INVALID
    assert lib.ne_b200_test_philox(None, 1, 0, 0, 4, None) == abi.ERR_INVALID


def test_missing_library_raises(tmp_path):
    with pytest.raises(abi.NarvalB200Error):
        abi.load_library(str(tmp_path / "nope.so"))


def test_product_package_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference legs may touch oracle/."""
    pkg = os.path.join(ROOT, "narvalengine_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "refclient" not in text and "narval_ref" not in text and "neref_" not in text, os.path.join(dirpath, f)
    assert "oracle" not in sys.modules
