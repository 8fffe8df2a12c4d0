"""In-tree native builds (no JIT cache: the .so files travel with the repo snapshot).

  libcsvsynth.so      host-only synthetic-input generator (g++)
  libcontextsv_b200.so  CUDA kernels + C ABI (nvcc, sm_100a only)

`python -m contextsv_b200.build` builds both; `__graft_entry__.build()` calls build_all().
"""
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INC = os.path.join(ROOT, "include")
LIB_CUDA = os.path.join(PKG, "libcontextsv_b200.so")
LIB_SYNTH = os.path.join(PKG, "libcsvsynth.so")

CUDA_SOURCES = ["capi.cu", "prep.cu", "walk.cu", "depth_tiles.cu", "radix_sort.cu", "sigs.cu", "dbscan1d.cu", "dbscan2d.cu", "windows.cu", "records.cu", "fetch.cu"]
HOST_SOURCES = ["widen.cpp"]          # plain g++ (AVX2 selected at run time)
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-Xptxas", "-v",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd, log=None):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        sys.stderr.write(p.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return p.stdout


def build_synth(force=False):
    src = [os.path.join(CSRC, "synth.cpp"), os.path.join(INC, "contextsv_b200.h")]
    if not force and _newer(LIB_SYNTH, src):
        return LIB_SYNTH
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-I" + INC, "-o", LIB_SYNTH, src[0], "-lpthread"])
    return LIB_SYNTH


def cuda_sources():
    return [os.path.join(CSRC, s) for s in CUDA_SOURCES + HOST_SOURCES]


def build_cuda(force=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = cuda_sources()
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [os.path.join(INC, "contextsv_b200.h")]
    if not force and _newer(LIB_CUDA, deps):
        return LIB_CUDA
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and %s is missing or stale" % LIB_CUDA)
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if s.endswith(".cu"):
            cmd = [nvcc] + NVCC_FLAGS + ["-I" + INC, "-I" + CSRC, "-c", s, "-o", o]
        else:
            o = os.path.join(objdir, os.path.basename(s)[:-4] + ".o"); objs[-1] = o
            cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-Wall", "-I" + INC, "-c", s, "-o", o]
        procs.append((cmd, o, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    ok = True
    for cmd, o, p in procs:
        out, _ = p.communicate()
        with open(o + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + out)
        if p.returncode != 0:
            sys.stderr.write(out)
            ok = False
    if not ok:
        raise RuntimeError("nvcc failed")
    _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_CUDA] + objs + ["-lcudart"])
    return LIB_CUDA


def build_oracle():
    """Test infrastructure: the C restatement, and (only where /root/reference exists) oracle/_ref."""
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    if os.path.exists("/root/reference/src/sv_caller.cpp"):
        _run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "ref"])
        if os.path.exists(LIB_CUDA):
            _run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "dropin"])


def build_all(force=False):
    build_synth(force)
    build_cuda(force)
    build_oracle()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built", LIB_CUDA, LIB_SYNTH)
