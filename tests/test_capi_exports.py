"""The C-ABI library loads and exports every symbol include/bunmpc.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bunmpc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bunmpc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from bunmpc_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libbunmpc.so not built: python __graft_entry__.py build"
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 16
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/bunmpc.h but not exported"
    assert set(syms) == set(_lib.EXPORTS), set(syms) ^ set(_lib.EXPORTS)


def test_version_and_defaults_without_gpu():
    from bunmpc_b200 import _lib
    L = _lib.lib()
    assert L.bunmpc_version() == 100
    p = _lib.Params()
    L.bunmpc_default_params(ctypes.byref(p))
    assert (p.max_outer, p.max_inner, p.tol, p.exit_tol, p.beta, p.mu, p.arith, p.slice_outer) == (100, 150, 1e-5, 1e-3, 1.5, 1.0, 0, 0)


def test_product_code_never_touches_the_oracle():
    """The shipped package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "bunmpc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r'#\s*include\s*[<"][^>"]*oracle', txt), f
                assert not re.search(r'^\s*(from|import)\s+oracle\b', txt, flags=re.M), f
                assert "libbicon_oracle" not in txt and "dlopen" not in txt, f


def test_multi_gpu_entry_points_reject_bad_arguments_without_gpu():
    """The argument checks of the job-counter / peer-result entry points come before any CUDA call."""
    from bunmpc_b200 import _lib
    L = _lib.lib()
    p, buf = ctypes.c_void_p(), ctypes.create_string_buffer(64)
    assert L.bunmpc_set_job_counter(None, None, 0) == _lib.ERR_ARG
    assert L.bunmpc_set_peer_results(None, 0, None) == _lib.ERR_ARG
    assert L.bunmpc_job_counter_create(0, None, buf) == _lib.ERR_ARG
    assert L.bunmpc_job_counter_open(0, None, ctypes.byref(p)) == _lib.ERR_ARG
    assert L.bunmpc_peer_buffer_create(0, 0, ctypes.byref(p), buf) == _lib.ERR_ARG      # zero bytes
    assert L.bunmpc_peer_buffer_open(0, buf.raw, None) == _lib.ERR_ARG
    assert L.bunmpc_peer_buffer_release(None, 1) == _lib.OK and L.bunmpc_job_counter_release(None, 0) == _lib.OK
    assert b"null" in L.bunmpc_last_error() or b"bad" in L.bunmpc_last_error()
