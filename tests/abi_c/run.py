"""Builds tests/abi_c/abi_smoke.c with gcc as strict C99 against include/onsas_cuda.h + libonsas_cuda.so and runs it."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def build_and_run():
    import onsas_jl_b200 as ob
    ob.build()
    libdir = os.path.dirname(ob._lib.SO_PATH)
    exe = os.path.join(tempfile.mkdtemp(prefix="onsas_abi_"), "abi_smoke")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           "-o", exe, os.path.join(HERE, "abi_smoke.c"), "-L" + libdir, "-lonsas_cuda", "-lm", "-Wl,-rpath," + libdir])
    return subprocess.run([exe], capture_output=True, text=True, timeout=120)
