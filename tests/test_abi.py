"""CPU tests of the drop-in boundary: libapc.so loads, exports every symbol that
include/*.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(apch?_[a-z0-9_]+)\s*\(", text)))


def test_headers_declare_something():
    assert len(declared_symbols("apc.h")) >= 20
    assert len(declared_symbols("apc_host.h")) >= 15


@pytest.mark.parametrize("header", ["apc.h", "apc_host.h"])
def test_every_declared_symbol_is_exported(built, header):
    from approx_counter_b200 import LIB_PATH
    lib = C.CDLL(LIB_PATH)
    for name in declared_symbols(header):
        assert hasattr(lib, name), f"{name} declared in include/{header} but not exported by libapc.so"


def test_binding_tables_cover_the_headers(built):
    from approx_counter_b200 import _lib, host
    assert sorted(_lib.SYMBOLS) == declared_symbols("apc.h")
    assert sorted(host.HOST_SYMBOLS) == declared_symbols("apc_host.h")


def test_headers_are_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "apc.h"\n#include "apc_host.h"\nint main(void){return APC_VERSION==2?0:1;}\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-c", str(src), "-o", str(tmp_path / "t.o")], check=True)


def test_no_cpu_fallback(built):
    """Without a CUDA device the library refuses to create a context."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from approx_counter_b200 import ApcError, ApproxCounter, load
    lib = load()
    assert lib.apc_version() == 2
    assert lib.apc_device_count() <= 0
    with pytest.raises(ApcError) as e:
        ApproxCounter(0)
    assert "NO_DEVICE" in str(e.value)
    assert lib.apc_strerror(-3) == b"no usable CUDA device"


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under approx_counter_b200/ may reference it."""
    pkg = os.path.join(ROOT, "approx_counter_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for line in text.splitlines():
                    code = line.split("//")[0].split("#")[0] if not f.endswith("Makefile") else line
                    assert "oracle" not in code.lower(), f"{os.path.join(dirpath, f)} uses the oracle: {line}"


def test_sass_is_sm100a(built):
    from approx_counter_b200 import LIB_PATH
    out = subprocess.run(["cuobjdump", "-lelf", LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_reference_shim_compiles_as_cpp14(built, tmp_path):
    """The header a reference maintainer would include builds without CUDA or SeqAn headers."""
    libdir = os.path.join(ROOT, "approx_counter_b200", "csrc")
    subprocess.run(["g++", "-std=c++14", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    "-I", os.path.join(ROOT, "integration"), os.path.join(ROOT, "integration", "shim_demo.cpp"),
                    "-o", str(tmp_path / "shim_demo"), "-L", libdir, "-lapc", f"-Wl,-rpath,{libdir}"], check=True)
    p = subprocess.run([str(tmp_path / "shim_demo")], capture_output=True)
    assert p.returncode == 2  # usage
