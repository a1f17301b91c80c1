"""CPU test: the nth_element replay used by the CUDA selection kernel (iv_slam_b200/csrc/introselect.h) produces the
same permutation as the real libstdc++ std::nth_element (what cv::KeyPointsFilter::retainBest runs)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def _build(tmp_path):
    so = os.path.join(str(tmp_path), "introselect_check.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, os.path.join(HERE, "native", "introselect_check.cpp")])
    L = ctypes.CDLL(so)
    L.introselect_mismatches.restype = ctypes.c_long
    L.heapselect_mismatches.restype = ctypes.c_long
    return L


def test_introselect_permutation_matches_libstdcxx(tmp_path):
    L = _build(tmp_path)
    assert L.introselect_mismatches(100000, 60, 8, 1, 0) == 0        # FAST-cell sized lists, heavy ties
    assert L.introselect_mismatches(5000, 3000, 30, 2, 0) == 0       # level sized lists
    assert L.introselect_mismatches(5000, 3000, 300, 3, 1) == 0      # sawtooth
    assert L.introselect_mismatches(5000, 3000, 300, 4, 2) == 0
    assert L.introselect_mismatches(5000, 500, 1, 5, 0) == 0         # all equal
    assert L.heapselect_mismatches(20000, 300, 10, 7) == 0           # depth-limit fallback branch
