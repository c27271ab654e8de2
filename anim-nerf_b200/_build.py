"""In-tree build of the C-ABI library (nvcc, sm_100a only) and of the oracle's C checker."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("AN_LIB_PATH") or os.path.join(HERE, "libanimnerf_b200.so")   # AN_LIB_PATH: A/B kernel variants (dev only)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
NVCC_FLAGS += os.environ.get("AN_NVCC_EXTRA", "").split()      # e.g. -DAN_FWD_SPLIT=0 (variant builds, dev only)
if os.environ.get("AN_MLP_TRACE"):      # debug timeline build (tools/trace_mlp.py); never the shipped library
    NVCC_FLAGS.append("-DAN_MLP_TRACE")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_lib(force=False, verbose=False):
    """Compile every csrc/*.cu into libanimnerf_b200.so (object files under build/)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "animnerf_b200.h")]
    objdir = os.path.join(ROOT, "build", "obj" + ("_" + os.path.basename(LIB) if os.environ.get("AN_LIB_PATH") else ""))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer(o, [s] + hdrs):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed on %s" % s)
    if force or procs or _newer(LIB, objs):
        subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart"])
    return LIB


def build_oracle():
    """Compile oracle/knn_oracle.c (test infrastructure; building the checker is not using it)."""
    src = os.path.join(ROOT, "oracle", "knn_oracle.c")
    out = os.path.join(ROOT, "oracle", "libknn_oracle.so")
    if _newer(out, [src]):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC", src, "-o", out, "-lm"])
    return out


def build_testlib():
    """Compile tests/csrc/*.cu (on-device fp32 reference kernels used by the GPU tests only) into
    tests/libanimnerf_b200_ref.so -- test infrastructure, kept out of the product library."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = sorted(glob.glob(os.path.join(ROOT, "tests", "csrc", "*.cu")))
    out = os.path.join(ROOT, "tests", "libanimnerf_b200_ref.so")
    hdrs = sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [os.path.join(ROOT, "include", "animnerf_b200.h")]
    if srcs and _newer(out, srcs + hdrs):
        subprocess.check_call([nvcc] + NVCC_FLAGS + ["-shared", "-o", out] + srcs + ["-lcudart"])
    return out


if __name__ == "__main__":
    build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_oracle()
    build_testlib()
    print(LIB)
