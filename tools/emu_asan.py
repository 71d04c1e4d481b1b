"""CPU memcheck of the SIMT kernel sources: builds the host-fiber emulation of the kernels (tests/emu) with
-fsanitize=address and runs tests/test_emu_kernels.py under it.  Every global / shared access of the real .cu sources
is then bounds-checked against the torch CPU allocations (reports carry the .cu file:line).  No GPU needed.

    python tools/emu_asan.py [pytest -k expression]
"""
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
import build_emu as b  # noqa: E402

OUT = "/tmp/pcl_emu_asan"


def build():
    os.makedirs(OUT, exist_ok=True)
    cpps = []
    for name in b.SOURCES:
        out = os.path.join(OUT, name[:-3] + ".emu.cpp")
        with open(out, "w") as f:
            f.write(f'#line 1 "{os.path.join(b.CSRC, name)}"\n' + b.rewrite(open(os.path.join(b.CSRC, name)).read()))
        cpps.append(out)
    cpps.append(os.path.join(b.HERE, "emu_tc_stubs.cpp"))
    lib = os.path.join(OUT, "libpcl_emu.so")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-ffp-contract=off", "-fno-strict-aliasing", "-w",
           "-fsanitize=address", "-fno-omit-frame-pointer", "-I", os.path.join(b.HERE, "shim"), "-I", b.CSRC,
           "-I", os.path.join(b.ROOT, "include"), "-x", "c++"] + cpps + ["-o", lib]
    subprocess.run(cmd, check=True)
    return lib


RUNNER = """
import sys
sys.path.insert(0, {emu!r})
import build_emu
build_emu.LIB = {lib!r}
build_emu.build = lambda force=False: {lib!r}
import pytest
sys.exit(pytest.main([{test!r}, "-q", "-p", "no:cacheprovider", "-k", {k!r}]))
"""

if __name__ == "__main__":
    lib = build()
    k = sys.argv[1] if len(sys.argv) > 1 else "not refuses"
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=asan,
               ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1:verify_asan_link_order=0")
    code = RUNNER.format(emu=os.path.join(ROOT, "tests", "emu"), lib=lib, test=os.path.join(ROOT, "tests", "test_emu_kernels.py"), k=k)
    sys.exit(subprocess.run([sys.executable, "-c", code], env=env, cwd=ROOT).returncode)
