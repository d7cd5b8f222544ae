"""CPU-side checks of the drop-in boundary: the shared library builds, loads, and exports every symbol
include/vslam_b200.h declares. No compute call is made (there is no GPU in the dev container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vslam_b200 import lib as vl
    if not os.path.exists(vl.LIB_PATH):
        vl.build_library()
    return ctypes.CDLL(vl.LIB_PATH)


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "vslam_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in include/vslam_b200.h but not exported: {missing}"


def test_binding_lists_every_symbol():
    from vslam_b200 import lib as vl
    assert sorted(vl.EXPORTS) == declared_symbols()


def test_struct_layout_matches_header():
    from vslam_b200 import lib as vl
    assert ctypes.sizeof(vl.PairResult) == 60 and ctypes.sizeof(vl.PairParams) == 24
    assert vl.PAIR_RESULT_DTYPE.fields["F"][1] == 24


def test_no_cpu_fallback_without_gpu(lib):
    """Without a CUDA device vb_create must fail loudly (VB_ERR_CUDA), never fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    lib.vb_create.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    rc = lib.vb_create(0, ctypes.byref(h))
    lib.vb_last_error.restype = ctypes.c_char_p
    assert rc == 2 and b"no CPU fallback" in lib.vb_last_error()


def test_product_sources_never_touch_oracle():
    """The oracle is the checker: nothing under vslam_b200/ or include/ may include, link or import it."""
    bad = []
    for base in ("vslam_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            if "build" in dp.split(os.sep) or "__pycache__" in dp:
                continue
            for f in fs:
                if f.endswith((".cu", ".cuh", ".cpp", ".h", ".py", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    if re.search(r"vb_oracle|oracle_lib|liboracle|vbo_|libvbref", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad
