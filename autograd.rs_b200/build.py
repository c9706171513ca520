"""Build recipe for the native parts (called by __graft_entry__.build()).

  libagrs_b200.so   : CUDA kernels + C ABI (include/agrs.h), nvcc, sm_100a only
  libagrs_engine.so : C++ host engine mirroring the reference's Rust API (include/agrs_engine.h),
                      plain g++ -- it includes no CUDA header and reaches the device only through agrs.h
Both are built in-tree (next to this file) so they travel with the repo snapshot to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
ENG = os.path.join(HERE, "engine")
INC = os.path.join(ROOT, "include")
LIB_KERNELS = os.path.join(HERE, "libagrs_b200.so")
LIB_ENGINE = os.path.join(HERE, "libagrs_engine.so")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr", "-I", INC, "-I", CSRC]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(cmd), flush=True)
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd[:3]) + " ...")
    if verbose and r.stdout.strip():
        print(r.stdout)


def build_kernels(force=False, verbose=False, extra=()):
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + [os.path.join(INC, "agrs.h")]
    if not force and not _newer(LIB_KERNELS, deps):
        return LIB_KERNELS
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + deps[len(srcs):]):
            cmd = [NVCC] + NVCC_FLAGS + list(extra) + ["-c", s, "-o", o]
            if verbose:
                print("+", " ".join(cmd), flush=True)
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on " + cmd[-3])
        if verbose and out.strip():
            print(out)
    # cudart is linked statically (nvcc default); no link-time dependency on libcuda (driver entry points are resolved at run time)
    _run([NVCC, "-shared", "-o", LIB_KERNELS] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"], verbose)
    return LIB_KERNELS


def build_engine(force=False, verbose=False):
    if not os.path.isdir(ENG):
        return None
    srcs = sorted(os.path.join(ENG, f) for f in os.listdir(ENG) if f.endswith(".cpp"))
    if not srcs:
        return None
    deps = srcs + [os.path.join(ENG, f) for f in os.listdir(ENG) if f.endswith((".hpp", ".h"))] + \
        [os.path.join(INC, f) for f in os.listdir(INC)] + [LIB_KERNELS]
    if not force and not _newer(LIB_ENGINE, deps):
        return LIB_ENGINE
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-I", INC, "-I", ENG, "-o", LIB_ENGINE] + srcs + \
        ["-L", HERE, "-lagrs_b200", "-Wl,-rpath,$ORIGIN"]
    _run(cmd, verbose)
    return LIB_ENGINE


def build_all(force=False, verbose=False):
    k = build_kernels(force, verbose)
    e = build_engine(force, verbose)
    return k, e


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
