"""Drop-in latency of one stereo frame through the C++ shim (two std::threads + matcher), with and without graph mode."""
import ctypes as C, os, sys, subprocess, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from iv_slam_b200 import api, synthetic as S

def build():
    api.lib()
    so = os.path.join(tempfile.mkdtemp(), "shim_check.so")
    subprocess.check_call(["g++", "-O2", "-std=c++14", "-fPIC", "-shared", "-pthread", "-o", so,
                           os.path.join(ROOT, "tests", "native", "shim_check.cpp"), os.path.join(ROOT, "shim", "ORBextractor.cc"),
                           os.path.join(ROOT, "shim", "Frame_ComputeStereoMatches.cc"), "-I", os.path.join(ROOT, "tests", "fake_opencv"),
                           "-I", os.path.join(ROOT, "shim"), "-I", os.path.join(ROOT, "include"), "-L", os.path.dirname(api.LIB_PATH),
                           "-livslam_gpu", "-Wl,-rpath," + os.path.dirname(api.LIB_PATH)])
    L = C.CDLL(so)
    L.shim_frame_latency_ms.restype = C.c_double
    return L

def measure(iters=200):
    L = build()
    left, right = S.make_stereo_pair(1241, 376, 0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    out = {}
    for g, name in ((0, "plain"), (1, "graph"), (3, "graph_pinned_input"), (5, "one_call"), (7, "one_call_pinned_input")):
        out[name] = L.shim_frame_latency_ms(p(left), p(right), 1241, 376, 2000, 20, 7, C.c_float(386.1448), C.c_float(718.856), iters, g)
    return out

if __name__ == "__main__":
    print(measure())
