"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly what include/anncur_b200.h declares."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from anncur_b200.csrc import build
    build.build()
    from anncur_b200 import _lib
    return _lib.load()


def _declared():
    src = open(os.path.join(ROOT, "include", "anncur_b200.h")).read()
    return sorted(set(re.findall(r"ANNCUR_API[^;(]*?\b(anncur_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from anncur_b200 import _lib
    declared = _declared()
    assert len(declared) >= 19
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (anncur_[a-z0-9_]+)", out)))
    assert exported == declared
    assert sorted(_lib.PROTOTYPES) == declared          # the ctypes binding covers the whole header
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_size_queries(lib):
    assert lib.anncur_abi_version() == 1
    # host-only queries (no device work)
    assert lib.anncur_pinv_workspace_bytes(2000, 500) >= 8 * (2000 * 500 + 500 * 500 + 500)
    assert lib.anncur_packed_items_bytes(100000, 500, 0) >= 2 * 16 * 100000 * 64
    assert lib.anncur_packed_items_bytes(100000, 500, 1) >= 16 * 100000 * 64
    assert lib.anncur_packed_items_bytes(0, 500, 0) == 256
    # kind F32R: both planes (k_i + 1 bound slot, padded to 64) plus the item-major fp32 copy: >= 8 bytes per element
    assert lib.anncur_packed_items_bytes(100000, 500, 2) >= 8 * 500 * 100000
    assert lib.anncur_packed_items_bytes(100000, 512, 2) > lib.anncur_packed_items_bytes(100000, 511, 2)   # the slot opens a k-block
    assert lib.anncur_score_dense_workspace_bytes(4096, 100000, 500, 2) >= 2 * 2 * 4096 * 512


def test_argument_errors_are_reported_without_a_gpu(lib):
    from anncur_b200 import _lib
    rc = lib.anncur_merge_topk(None, None, 4, 8, 0, None, None, None)
    assert rc == _lib.E_INVALID and b"k = 0" in lib.anncur_last_error()
    rc = lib.anncur_topk_rows_f32(None, 0, 3, 10, 5000, 0, None, None, None)
    assert rc == _lib.E_INVALID
    assert lib.anncur_gemm_f32(None, 0, None, 0, None, 0, 0, 5, 3, None) == 0      # empty problem is a no-op
    # the kind checks of the packed index come before any device work (pointers below are never dereferenced on the host)
    import ctypes as C
    p = C.c_void_p(256)
    rc = lib.anncur_pack_items(p, 100, 100, 16, 7, p, p, None)
    assert rc == _lib.E_INVALID and b"kind" in lib.anncur_last_error()
    rc = lib.anncur_pack_items(p, 100, 100, 9000, 2, p, p, None)
    assert rc == _lib.E_UNSUPPORTED and b"k_dim" in lib.anncur_last_error()
    rc = lib.anncur_score_topk(p, 16, 4, p, p, 100, 16, 0, 5000, 0, p, p, p, 1 << 20, None)
    assert rc == _lib.E_INVALID and b"k = 5000" in lib.anncur_last_error()
    rc = lib.anncur_score_dense(p, 16, 4, p, p, 100, 16, 1, p, 100, p, 1 << 20, None)                 # bf16 has no fp32-grade planes
    assert rc == _lib.E_INVALID and b"kind" in lib.anncur_last_error()


def test_header_is_plain_c_and_links(lib, tmp_path):
    """The boundary is a C ABI: the header compiles as C99 (no C++-isms, no torch types) and a C program links and calls it."""
    from anncur_b200 import _lib
    src = tmp_path / "use_abi.c"
    src.write_text(
        '#include "anncur_b200.h"\n#include <stdio.h>\n'
        'int main(void) {\n'
        '    size_t b = anncur_packed_items_bytes(1000, 500, ANNCUR_KIND_F32R);\n'
        '    int rc = anncur_merge_topk(NULL, NULL, 4, 8, 0, NULL, NULL, NULL);\n'
        '    printf("%d %zu %d %s\\n", anncur_abi_version(), b, rc, anncur_last_error());\n'
        '    return (anncur_abi_version() == ANNCUR_ABI_VERSION && rc == ANNCUR_E_INVALID) ? 0 : 1;\n}\n')
    exe = tmp_path / "use_abi"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lanncur_b200", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.split()[0] == "1" and int(r.stdout.split()[1]) >= 8 * 500 * 1000


def test_built_for_sm100a_with_tcgen05_and_tma():
    from anncur_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):          # tcgen05.mma, TMA load, tcgen05.ld
        assert mnemonic in sass, mnemonic


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "anncur_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f
