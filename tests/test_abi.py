"""CPU-side checks of the C-ABI boundary: the library loads without a GPU and exports every symbol that
include/svt_av1_b200.h declares."""
import ctypes as C
import os
import re

import svtb200 as sb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    """Every function the header declares, after macro expansion by the C preprocessor."""
    import subprocess
    text = subprocess.check_output(["gcc", "-E", "-P", os.path.join(ROOT, "include", "svt_av1_b200.h")]).decode()
    return sorted(set(re.findall(r'visibility\("default"\)\)\)[^;(]*?\b(\w+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    lib = sb.load()
    names = declared_symbols()
    assert len(names) > 60
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_sizes_match_header():
    # sizes computed by a C compiler on the header itself
    import subprocess, tempfile
    src = '#include "svt_av1_b200.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(SvtB200Plane), sizeof(SvtB200MeParams), sizeof(SvtB200MePlanes), sizeof(SvtB200HmeResult), sizeof(SvtB200MeOutputs));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        got = [int(x) for x in subprocess.check_output([os.path.join(d, "t")]).split()]
    want = [C.sizeof(sb.Plane), C.sizeof(sb.MeParams), C.sizeof(sb.MePlanes), C.sizeof(sb.HmeResult), C.sizeof(sb.MeOutputs)]
    assert got == want


def test_no_gpu_is_an_error_not_a_fallback():
    lib = sb.load()
    n = lib.svt_b200_device_count()
    if n <= 0:
        # without a device the picture-level entry must fail loudly
        p = sb.preset8_me_params(128, 128)
        rc = lib.svt_b200_me_picture(C.byref(p), None, None, None, None, None)
        assert rc != 0
