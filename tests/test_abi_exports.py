"""CPU tests of the drop-in boundary: the shared library loads and exports every symbol that
include/liloc_gpu.h declares; no compute is called (there is no GPU here)."""
import ctypes
import os
import re

import pytest

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


def test_struct_layouts_match_header(tmp_path):
    """Every ctypes mirror in liloc_b200/abi.py against sizeof / offsetof of the real header, compiled here with gcc (C, not C++:
    the header is the C ABI)."""
    import shutil
    import subprocess
    from liloc_b200 import abi
    assert ctypes.sizeof(abi.S2mOpts) == 8 * 4
    assert ctypes.sizeof(abi.S2mResult) == 8 * 4 + 72 * 4
    assert ctypes.sizeof(abi._Config) == 8 * 4
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not found")
    mirrors = {"liloc_s2m_opts": abi.S2mOpts, "liloc_s2m_result": abi.S2mResult, "liloc_config": abi._Config, "liloc_frame_in": abi.FrameIn,
               "liloc_s2m_extra": abi.S2mExtra, "liloc_timing": abi.Timing, "liloc_stats": abi.Stats, "liloc_ndt_opts": abi.NdtOpts,
               "liloc_ndt_stats": abi.NdtStats, "liloc_ndt_job": abi.NdtJob, "liloc_ndt_out": abi.NdtOut, "liloc_deskew_params": abi.DeskewParams}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "liloc_gpu.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  printf("liloc_point %zu\\n", sizeof(liloc_point));', '  printf("liloc_ring_time %zu\\n", sizeof(liloc_ring_time));', '  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    r = subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr            # also proves the header is plain C
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for cname, cls in mirrors.items():
        assert int(got[cname]) == ctypes.sizeof(cls), (cname, got[cname], ctypes.sizeof(cls))
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)
    assert int(got["liloc_point"]) == 16 and int(got["liloc_ring_time"]) == abi.RING_TIME.itemsize == 8


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
