"""Builds libzprize_b200.so (nvcc, sm_100a) in-tree.  `python build.py [--emu]`.

--emu builds tests/emu/libzprize_emu.so instead: the SAME sources compiled by g++ against the CPU
emulation layer in tests/emu (unit-test infrastructure only; never loaded by the product path).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["ntt.cu", "msm.cu", "poly.cu", "wiring.cu", "witness.cu", "prover.cu", "verifier.cu", "capi.cu"]
HEADERS = ["ptx_ops.cuh", "field.cuh", "curve.cuh", "common.cuh", "ntt.cuh", "msm.cuh", "msm_affine.cuh", "poly.cuh", "wiring.cuh", "gates.cuh",
           "host_math.hpp", "pairing.hpp", "transcript.hpp", "prover.cuh"]
LIB = os.path.join(HERE, "libzprize_b200.so")
EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "libzprize_emu.so")


def _newer_than(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(verbose=False):
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(ROOT, "include", "zprize_b200.h")]
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
             "-ccbin", "g++", "--expt-relaxed-constexpr"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    jobs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        if _newer_than(obj, deps):
            jobs.append([nvcc] + flags + ["-c", os.path.join(CSRC, src), "-o", obj])
    with ThreadPoolExecutor(max_workers=8) as ex:
        outs = list(ex.map(_run, jobs))
    if verbose:
        for o in outs:
            print(o)
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or not os.path.exists(LIB):
        _run([nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"])
    return LIB


def build_emu():
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(EMU_DIR, "cuda_emu.h"),
                                                                  os.path.join(EMU_DIR, "cuda_emu.cpp"),
                                                                  os.path.join(ROOT, "include", "zprize_b200.h")]
    if not _newer_than(EMU_LIB, deps):
        return EMU_LIB
    cxx = os.environ.get("ZP_CXX", "g++")
    objdir = os.path.join(EMU_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    flags = ["-std=c++17", "-O2", "-fPIC", "-DZP_EMU", "-include", os.path.join(EMU_DIR, "cuda_emu.h"), "-Wno-unused-function",
             "-Wno-attributes", "-Wno-unknown-pragmas"]
    jobs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        jobs.append([cxx] + flags + ["-x", "c++", "-c", os.path.join(CSRC, src), "-o", obj])
    jobs.append([cxx, "-std=c++17", "-O2", "-fPIC", "-c", os.path.join(EMU_DIR, "cuda_emu.cpp"), "-o",
                 os.path.join(objdir, "cuda_emu.o")])
    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(_run, jobs))
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES] + [os.path.join(objdir, "cuda_emu.o")]
    _run([cxx, "-shared", "-o", EMU_LIB] + objs)
    return EMU_LIB


if __name__ == "__main__":
    if "--emu" in sys.argv:
        print(build_emu())
    else:
        print(build(verbose="-v" in sys.argv))
