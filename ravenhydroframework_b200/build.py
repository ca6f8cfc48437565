"""In-tree build of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  librvn_gpu.so      sm_100a CUDA engine + C ABI (include/rvn_gpu.h); cudart linked statically so the
                     library loads (and reports "no device") on a CPU-only box
  libraven_b200.so   Raven-surface C++ host layer (parsers, CModel, flatten, MassEnergyBalance, BMI)
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # keep the reference's operation order: no FMA contraction
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-diag-suppress", "550",
]
HOST_SRC = ["rvh_util.cpp", "rvh_model.cpp", "rvh_process.cpp", "rvh_parse.cpp", "rvh_flatten.cpp", "rvh_solver.cpp",
            "rvh_capi.cpp", "rvh_bmi.cpp", "rvh_ensemble.cpp"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, cwd=None):
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build failed: " + " ".join(cmd))
    return r.stdout


def build_engine(force=False, verbose=False):
    csrc = os.path.join(HERE, "csrc")
    out = os.path.join(HERE, "librvn_gpu.so")
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(ROOT, "include", "rvn_gpu.h")]
    if force or _newer(out, deps):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I" + os.path.join(ROOT, "include"),
                                                                           os.path.join(csrc, "rvn_engine.cu"), "-o", out]
        log = _run(cmd)
        if verbose:
            print(log)
    return out


def build_host(force=False):
    hdir = os.path.join(HERE, "host")
    out = os.path.join(HERE, "libraven_b200.so")
    srcs = [os.path.join(hdir, f) for f in HOST_SRC if os.path.exists(os.path.join(hdir, f))]
    deps = [os.path.join(hdir, f) for f in os.listdir(hdir)] + [os.path.join(ROOT, "include", "rvn_gpu.h")]
    if force or _newer(out, deps):
        cmd = ["g++", "-std=c++14", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-I" + hdir] + srcs + \
              ["-o", out, "-L" + HERE, "-lrvn_gpu", "-Wl,-rpath,$ORIGIN"]
        _run(cmd)
    return out


def build_all(force=False, verbose=False):
    build_engine(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("ok")
