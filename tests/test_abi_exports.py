"""CPU tests of the drop-in boundary: the shared library loads and exports every symbol that
include/liloc_gpu.h declares; no compute is called (there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "liloc_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(liloc_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path_entry_points():
    names = _declared()
    for must in ("liloc_create", "liloc_destroy", "liloc_keyframe_put", "liloc_localmap_build", "liloc_scan_put",
                 "liloc_scan2map", "liloc_frame", "liloc_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from liloc_b200 import abi
    L = ctypes.CDLL(abi.lib_path)
    for name in _declared():
        assert hasattr(L, name), f"{name} declared in include/liloc_gpu.h but not exported"
    assert set(abi.EXPORTS) <= set(_declared())
    assert b"sm_100a" in abi.lib.liloc_version()


def test_library_is_sm100a_only():
    """The product is compiled for sm_100a and nothing else (no multi-arch fat binary)."""
    import shutil
    import subprocess
    from liloc_b200 import abi
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    out = subprocess.run([cuobjdump, "--list-elf", abi.lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_struct_layouts_match_header():
    from liloc_b200 import abi
    assert ctypes.sizeof(abi.S2mOpts) == 8 * 4
    assert ctypes.sizeof(abi.S2mResult) == 8 * 4 + 72 * 4
    assert ctypes.sizeof(abi._Config) == 8 * 4


def test_product_does_not_reference_the_oracle():
    """The product path must not import, include or link anything under oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "liloc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r'#include\s+"[^"]*oracle|import\s+oracle|from\s+oracle|liboracle|libliloc_oracle', txt) and f != "_build.py":
                    bad.append(os.path.join(base, f))
    assert not bad, bad
