"""The C-ABI boundary without a GPU: libvolta_b200.so loads, exports every entry point include/volta_b200.h declares,
and the ctypes binding (volta_b200/_lib.py) covers exactly that set.  No compute calls (no device here)."""
import ctypes
import os
import re

import pytest

from volta_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "volta_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"^\s*(?:int|const\s+char\s*\*)\s+(vb_\w+)\s*\(", text, re.M))


def test_header_declares_the_bound_symbols():
    names = _declared()
    assert len(names) >= 40
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build with `python -m volta_b200.csrc.build`"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in sorted(_declared()) if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_loads_and_reports_version_without_a_device():
    lib = _lib.load()
    assert lib.vb_version() >= 1000
    assert isinstance(_lib.last_error(), str)


def test_signatures_have_as_many_arguments_as_the_header():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % name, text, re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(argtypes), (name, n, len(argtypes))


def test_c_program_binds_the_library(tmp_path):
    """The boundary from C, as a cgo / JNI stub would bind it: the header is valid C99 (gcc -std=c99 -pedantic), a C
    program links against the shared library, reads the ABI version and gets argument errors + messages back without a
    device (no compute call)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    src = tmp_path / "bind.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "volta_b200.h"
int main(void) {
  if (vb_version() < 1000) return 1;
  /* null pointers: rejected before any CUDA call */
  if (vb_layernorm_fwd(NULL, NULL, NULL, NULL, NULL, NULL, 8, 768, 1e-12f, 1.0f, 0, NULL, NULL, NULL, 0, VB_BF16, NULL) != VB_ERR_ARG) return 2;
  if (strlen(vb_last_error()) == 0) return 3;
  if (vb_gemm(NULL, NULL) == 0) return 4;
  if (vb_attention_fwd(NULL, NULL) == 0) return 5;
  if (vb_bce_logits_fwd(NULL, 0, NULL, 0, 0, 0, NULL, 1.0f, NULL, VB_F32, NULL) != VB_ERR_ARG) return 6;
  printf("abi %d\n", vb_version());
  return 0;
}
''')
    exe = tmp_path / "bind"
    libdir = os.path.dirname(_lib.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe),
           "-L", libdir, "-lvolta_b200", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith("abi ")
