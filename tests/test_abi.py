"""The C-ABI library loads, exports every symbol include/artes_gpu.h declares, and the ctypes
mirrors of its structs have the C layout.  No compute calls (CPU box)."""
import ctypes as C
import os
import re
import subprocess

from artes_b200 import abi
from artes_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "artes_gpu.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(artes_gpu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    lib = C.CDLL(L.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 15
    for s in decl:
        assert hasattr(lib, s), f"{s} declared in artes_gpu.h but not exported"
    assert sorted(L.SYMBOLS) == decl


def test_abi_version_and_load():
    lib = L.load()
    assert lib.artes_gpu_abi_version() == abi.ABI_VERSION


def test_struct_layout_matches_c(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "artes_gpu.h"\n'
                    'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(artes_launch_t), offsetof(artes_launch_t, seed),'
                    ' offsetof(artes_launch_t, fstop), offsetof(artes_launch_t, y_max), sizeof(artes_stats_t),'
                    ' offsetof(artes_stats_t, kernel_ms));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    vals = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert vals == [C.sizeof(abi.Launch), abi.Launch.seed.offset, abi.Launch.fstop.offset, abi.Launch.y_max.offset,
                    C.sizeof(abi.Stats), abi.Stats.kernel_ms.offset]


def test_create_fails_loudly_without_gpu():
    """No CPU fallback: on a box without a CUDA device the product path must raise."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        L.GpuTransport((0,))
    except L.ArtesGpuError as e:
        assert "no CUDA device" in str(e) or "fallback" in str(e)
    else:
        raise AssertionError("GpuTransport() succeeded without a GPU")
