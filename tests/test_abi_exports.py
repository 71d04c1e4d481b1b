"""The C-ABI library loads and exports every symbol include/pcl.h declares (no compute calls: CPU-only)."""
import os
import re

import pytest

import contrastiveseg_b200 as cs
from contrastiveseg_b200 import _abi, build

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pcl.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pcl_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_loads():
    path = build.build_library()
    assert os.path.exists(path)
    lib = _abi.load()
    assert lib.pcl_version() == 100


def test_every_declared_symbol_is_exported_and_bound():
    lib = _abi.load(build_if_missing=True)
    syms = declared_symbols()
    assert len(syms) >= 20
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in pcl.h but not exported"
        assert name in _abi.SIGNATURES, f"{name} has no ctypes signature"
    for name in _abi.SIGNATURES:
        assert name in syms, f"{name} bound in _abi.py but not declared in pcl.h"


def test_status_strings():
    lib = _abi.load(build_if_missing=True)
    assert lib.pcl_strerror(0) == b"ok"
    assert b"argument" in lib.pcl_strerror(-1)


def test_struct_layout_matches_header():
    # sizes of the host structs as laid out by the C compiler rules (natural alignment)
    import ctypes as C
    assert C.sizeof(_abi.Geom) == 40
    assert C.sizeof(_abi.BankGeom) == 40
    assert C.sizeof(_abi.SelectSizes) == 48
    assert C.sizeof(_abi.SweepSizes) == 40
    # every host struct, against the sizes the compiled library reports
    lib = _abi.load(build_if_missing=True)
    for sid, st in enumerate(_abi.ABI_STRUCTS):
        assert lib.pcl_abi_sizeof(sid) == C.sizeof(st), st.__name__
    assert lib.pcl_abi_sizeof(len(_abi.ABI_STRUCTS)) == -1


def test_no_cpu_fallback():
    import torch
    with pytest.raises(_abi.PclError):
        cs.pixel_contrast_loss(torch.zeros(1, 32, 4, 4), torch.zeros(1, 8, 8, dtype=torch.long),
                               predict=torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(_abi.PclError):
        cs.l2_normalize(torch.zeros(1, 32, 4, 4))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "contrastiveseg_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "import oracle" not in src and "from oracle" not in src, f


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/pcl.h must compile as C99 with nothing but the standard headers."""
    import subprocess
    src = tmp_path / "chk.c"
    src.write_text('#include "pcl.h"\nint main(void) { pcl_geom g; (void)g; return pcl_version() > 0 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I",
                        os.path.join(ROOT, "include"), str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_client_links_and_runs(tmp_path):
    """examples/c_abi_sizes.c: a plain C program against include/pcl.h + libpcl_b200.so (host-only entry points)."""
    import subprocess
    lib = build.build_library()
    exe = tmp_path / "c_abi_sizes"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "examples", "c_abi_sizes.c"), "-L", os.path.dirname(lib), "-lpcl_b200",
                        f"-Wl,-rpath,{os.path.dirname(lib)}", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "selection scratch: keys 262144 u16" in r.stdout and "pcl 100" in r.stdout
