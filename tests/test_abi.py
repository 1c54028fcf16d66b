"""The C-ABI library loads without a GPU and exports exactly what include/rvb.h declares; the ctypes
signature table mirrors the header (argument counts and kinds)."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rvb.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|int64_t|const char\*)\s+(rvb_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",")]
        decls[m.group(2)] = [] if args == ["void"] else args
    return decls


@pytest.fixture(scope="module")
def lib():
    from reconvat_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def test_header_and_library_agree(lib):
    decls = _declared()
    assert len(decls) >= 17
    out = subprocess.run(["nm", "-D", "--defined-only", lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line and "rvb_" in line}
    assert exported == set(decls), (exported ^ set(decls))
    handle = lib.load()
    for name in decls:
        assert hasattr(handle, name)
    assert handle.rvb_abi_version() == lib.ABI_VERSION == int(re.search(r"RVB_ABI_VERSION (\d+)", open(HEADER).read()).group(1))


def test_ctypes_table_mirrors_header(lib):
    import ctypes
    decls = _declared()
    plumbing = {"rvb_abi_version", "rvb_last_error", "rvb_launch_count", "rvb_vat_stats_workspace_bytes",
                "rvb_parity_plane_len", "rvb_bn_splits", "rvb_bn_nhwc_workspace_bytes"}
    assert set(lib.SIGNATURES) == set(decls) - plumbing
    kinds = {ctypes.c_void_p: "ptr", ctypes.c_int: "int", ctypes.c_int64: "int64_t", ctypes.c_float: "float",
             ctypes.c_double: "double", ctypes.c_uint64: "uint64_t", ctypes.c_uint32: "uint32_t"}
    for name, argtypes in lib.SIGNATURES.items():
        args = decls[name]
        assert len(args) == len(argtypes), name
        for a, t in zip(args, argtypes):
            if "*" in a or a.startswith("rvb_stream_t"):
                assert kinds[t] == "ptr", (name, a)
            elif a.startswith("int64_t"):
                assert kinds[t] == "int64_t", (name, a)
            elif a.startswith("uint64_t") or a.startswith("uint32_t"):
                assert kinds[t] == a.split()[0], (name, a)
            elif a.startswith("int"):
                assert kinds[t] == "int", (name, a)
            elif a.startswith("float"):
                assert kinds[t] == "float", (name, a)
            elif a.startswith("double"):
                assert kinds[t] == "double", (name, a)
            else:
                raise AssertionError("unparsed argument %r of %s" % (a, name))
    # no torch / C++ types, only plain pointers and sizes in the declarations (comments stripped)
    code = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    assert "torch" not in code.lower() and "at::" not in code and "std::" not in code


def test_argument_errors_are_reported_not_thrown(lib):
    """Argument validation happens before any CUDA call, so it can be exercised without a GPU."""
    h = lib.load()
    assert h.rvb_vat_perturb(None, None, None, 4, 229, 1e-6, 1, None) == -1
    assert b"null pointer" in h.rvb_last_error()
    assert h.rvb_stft_gemm(1, 1, 1, 644, 500, 640, 1, 1, 2048, 2048, 0, 2.0, 1, 1024, None) == -1
    assert b"hop" in h.rvb_last_error()
    assert h.rvb_bce_mean(1, 1, 0, 1, 1, None) == -1
    assert h.rvb_launch_count() == 0


def test_missing_library_fails_loudly(lib, monkeypatch):
    monkeypatch.setattr(lib, "_lib", None)
    monkeypatch.setattr(lib, "LIB_PATH", "/nonexistent/librvb.so")
    with pytest.raises(ImportError, match="no CPU"):
        lib.load()
