"""CPU tests of the drop-in boundary: the C-ABI library builds, loads, exports every symbol include/ivslam_gpu.h
declares, and fails loudly (no CPU fallback) when there is no GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "ivslam_gpu.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ivg_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from iv_slam_b200 import api
    L = api.lib()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), "libivslam_gpu.so does not export %s" % n
    assert set(api.exported_symbols()) == set(names), set(api.exported_symbols()) ^ set(names)


def test_keypoint_record_layout_is_cv_keypoint():
    from iv_slam_b200 import api
    assert api.KP_DTYPE.itemsize == 28
    assert [api.KP_DTYPE.fields[f][1] for f in ("x", "y", "size", "angle", "response", "octave", "class_id")] == [0, 4, 8, 12, 16, 20, 24]


def test_strerror_and_argument_checks():
    from iv_slam_b200 import api
    L = api.lib()
    assert L.ivg_strerror(0) == b"ok"
    assert b"device" in L.ivg_strerror(-5)
    h = ctypes.c_void_p()
    assert L.ivg_extractor_create(ctypes.byref(h), 0, 0, 1.2, 8, 20, 7, 0) == -1      # nfeatures < 1
    assert L.ivg_extractor_create(ctypes.byref(h), 0, 1000, 1.2, 99, 20, 7, 0) == -1   # too many levels
    assert L.ivg_run_batch(None) == -6


def test_no_gpu_means_error_not_fallback():
    """On a box without a GPU the product must refuse to run rather than compute on the CPU."""
    import torch
    from iv_slam_b200 import api
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    with pytest.raises(api.IvgError) as e:
        api.ORBextractor(1000, 1.2, 8, 20, 7)
    assert e.value.status in (-5, -4)


def test_product_does_not_reference_the_oracle():
    """iv_slam_b200/ must never import, link or call oracle/ (it is test infrastructure)."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "iv_slam_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"\boracle\b", txt):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_header_is_valid_c99_and_cxx11(tmp_path):
    """include/ivslam_gpu.h is the boundary a C or C++ host binds: it must compile on its own in both languages."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "ivslam_gpu.h"\nint main(void) { ivg_extractor* h = 0; (void)h; return IVG_OK; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", inc, str(src)])
    subprocess.check_call(["g++", "-std=c++11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", "-I", inc, str(src)])
