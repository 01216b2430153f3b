"""CPU: the C-ABI library builds/loads and exports every symbol include/mage_b200.h declares
(no compute calls without a GPU)."""
import os
import re

from mage_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "mage_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mage_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mage_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert L.mage_abi_version() == _lib.ABI_VERSION
    assert L.mage_launch_count(None) == 0          # no handle, nothing launched
    import ctypes
    import torch
    if not torch.cuda.is_available():
        # the handle is the only way in, and it only exists on an sm_100 device: no CPU path to fall back to
        h = ctypes.c_void_p()
        assert L.mage_ctx_create(0, ctypes.byref(h)) != 0 and not h.value


def test_header_cites_reference_lines():
    text = open(os.path.join(ROOT, "include", "mage_b200.h")).read()
    assert text.count("mage_model.py:") >= 8 and text.count("vqvae_model.py:") >= 6


def test_product_never_imports_the_oracle():
    """oracle/ (and the oracle/_ref copy of the reference) is test / bench infrastructure: no product module may import it."""
    bad = []
    files = [os.path.join(ROOT, f) for f in ("main_mage.py", "dataload.py")]
    for d in ("mage_b200", "modules", "utils"):
        for dp, _, fs in os.walk(os.path.join(ROOT, d)):
            files += [os.path.join(dp, f) for f in fs if f.endswith(".py")]
    for f in files:
        for ln in open(f):
            if re.match(r"\s*(from|import)\s+oracle\b", ln) or "oracle/_ref" in ln or "ref_shims" in ln:
                bad.append((f, ln.strip()))
    assert not bad, bad


def test_header_is_plain_c_and_a_c_program_can_bind_it(tmp_path):
    """include/mage_b200.h is the boundary a maintainer binds from any language: it must parse as strict C99 (no torch / CUDA types)
    and as C++, and a C program that takes the address of every declared entry point must link against the library."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("no gcc")
    hdr = os.path.join(ROOT, "include", "mage_b200.h")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr], check=True)
    subprocess.run(["g++", "-std=c++11", "-Werror", "-fsyntax-only", "-x", "c++", hdr], check=True)
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    src = tmp_path / "bind.c"
    names = _declared()
    src.write_text('#include "mage_b200.h"\n#include <stdio.h>\nint main(void) {\n  const void* f[] = {'
                   + ", ".join(f"(const void*){n}" for n in names)
                   + '};\n  printf("%d %d\\n", (int)(sizeof f / sizeof f[0]), mage_abi_version());\n  return 0;\n}\n')
    exe = tmp_path / "bind"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wno-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", libdir,
                    "-l:" + os.path.basename(_lib.LIB_PATH), "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert int(out[0]) == len(names) and int(out[1]) == _lib.ABI_VERSION
