"""Builds surfcascade_b200/libsurfcascade_b200.so in-tree: hand-written sm_100a kernels + the C-ABI + the kept host classes."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsurfcascade_b200.so")
CLI = os.path.join(HERE, "ObjDetector")
LIB_CHECKED = os.path.join(HERE, "libsurfcascade_b200_checked.so")   # -DSC_CHECKED: range-tested gathers (tests/test_gpu_checked_build.py)
CLASS_TEST = os.path.join(HERE, "class_detect")   # tests/cpp/class_detect.cpp: the reference's detect branch over the kept classes
CU = ["csrc/sc_capi.cu"]
CPP = ["host/sc_host.cpp", "host/cfgfile.cpp", "host/classes.cpp", "host/Model.cpp", "host/DenseSURFFeatureExtractor.cpp"]
HEADERS = ["csrc/sc_kernels.cuh", "csrc/sc_plan.h", "csrc/sc_comm.inc", "host/sc_host.h", "host/sc_access.h", "host/cfgfile.h", "host/Model.h", "host/cvcompat.h",
           "host/CascadeClassifier/CascadeClassifier.h", "host/CascadeClassifier/GentleAdaboost.h", "host/CascadeClassifier/LogisticRegression.h",
           "host/CascadeClassifier/StageClassifier.h", "host/FeatureExtractors/DenseSURFFeatureExtractor.h", "../include/surfcascade.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-I", os.path.join(HERE, "host"), "-I", os.path.join(HERE, "..", "include")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in CU + CPP + HEADERS + ["build.py"])


def build_checked() -> str:
    subprocess.check_call([_nvcc()] + NVCC_FLAGS + ["-DSC_CHECKED", "-shared", "-o", LIB_CHECKED] + [os.path.join(HERE, f) for f in CU + CPP])
    return LIB_CHECKED


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    extra = os.environ.get("SC_EXTRA_NVCC", "").split()  # tuning experiments, e.g. -DSC_STAGE0_MIN_CTAS=4
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", LIB] + [os.path.join(HERE, f) for f in CU + CPP]
    subprocess.check_call(cmd)
    if os.path.exists(os.path.join(HERE, "host", "ObjDetector.cpp")):
        subprocess.check_call([_nvcc()] + NVCC_FLAGS + ["-o", CLI, os.path.join(HERE, "host", "ObjDetector.cpp"), "-L", HERE, "-lsurfcascade_b200",
                               "-Xlinker", "-rpath,$ORIGIN"])
    src = os.path.join(HERE, "..", "tests", "cpp", "class_detect.cpp")
    if os.path.exists(src):
        subprocess.check_call([_nvcc()] + NVCC_FLAGS + ["-o", CLASS_TEST, src, "-L", HERE, "-lsurfcascade_b200", "-Xlinker", "-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    if "--checked" in sys.argv:
        print(build_checked())
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
