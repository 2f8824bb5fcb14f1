"""CPU: the C-ABI library loads and exports every symbol include/aprb200.h declares (no compute calls)."""
import ctypes
import os
import re

from apr_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "aprb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = re.findall(r"\b(?:int|size_t|long long|const char\*)\s+(aprb_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S)
    return {name: [a.strip() for a in args.split(",") if a.strip() and a.strip() != "void"] for name, args in decls}


def test_library_exports_every_declared_symbol():
    fns = _header_functions()
    assert len(fns) >= 14
    assert os.path.exists(_native.LIB_PATH), "libaprb200.so not built (run __graft_entry__.build())"
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in fns:
        assert hasattr(lib, name), f"{name} declared in aprb200.h but not exported"


def test_python_binding_matches_header_arity():
    fns = _header_functions()
    assert set(fns) == set(_native.SIGNATURES), set(fns) ^ set(_native.SIGNATURES)
    for name, args in fns.items():
        assert len(args) == len(_native.SIGNATURES[name][1]), name


def test_version_and_error_string_without_gpu():
    lib = _native.lib()
    assert lib.aprb_version() >= 100
    assert isinstance(lib.aprb_last_error(), bytes)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "apr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle"
                assert "liboracle" not in text and "libapr_ref" not in text, f"{f} references the oracle libraries"
