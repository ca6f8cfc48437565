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


def _digest(sources, flags):
    """sha256 over the compiler command line and the contents of every source / header the library is built from."""
    import hashlib
    h = hashlib.sha256()
    h.update("\0".join(flags).encode())
    for p in sorted(sources):
        h.update(os.path.basename(p).encode() + b"\0")
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale(target, digest):
    """A library is current only if the digest stamped next to it (and embedded in it, see RVN_BUILD_DIGEST) is the one of the
    present sources and flags: mtimes are not trusted (the .so files travel between boxes)."""
    stamp = target + ".digest"
    if not os.path.exists(target) or not os.path.exists(stamp):
        return True
    return open(stamp).read().strip() != digest


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
    digest = _digest(deps, [os.path.basename(NVCC)] + NVCC_FLAGS)
    if force or _stale(out, digest):
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-DRVN_BUILD_DIGEST=\"%s\"" % digest, "-I" + os.path.join(ROOT, "include"),
                                                                           os.path.join(csrc, "rvn_engine.cu"), "-o", out]
        log = _run(cmd)
        open(out + ".digest", "w").write(digest + "\n")
        if verbose:
            print(log)
    return out


def build_host(force=False):
    hdir = os.path.join(HERE, "host")
    out = os.path.join(HERE, "libraven_b200.so")
    srcs = [os.path.join(hdir, f) for f in HOST_SRC if os.path.exists(os.path.join(hdir, f))]
    deps = [os.path.join(hdir, f) for f in os.listdir(hdir)] + [os.path.join(ROOT, "include", "rvn_gpu.h")]
    flags = ["g++", "-std=c++14", "-O2", "-fPIC", "-shared"]
    digest = _digest(deps, flags)
    if force or _stale(out, digest):
        cmd = flags + ["-DRVH_BUILD_DIGEST=\"%s\"" % digest, "-I" + os.path.join(ROOT, "include"), "-I" + hdir] + srcs + \
              ["-o", out, "-L" + HERE, "-lrvn_gpu", "-Wl,-rpath,$ORIGIN"]
        _run(cmd)
        open(out + ".digest", "w").write(digest + "\n")
    return out


def build_all(force=False, verbose=False):
    build_engine(force, verbose)
    build_host(force)


def source_digests():
    """{library: digest of the sources and flags it would be built from now} - compared with rvn_gpu_build_digest() /
    rvh_build_digest() of the loaded libraries by tests/test_abi_and_host.py."""
    csrc, hdir = os.path.join(HERE, "csrc"), os.path.join(HERE, "host")
    hdr = [os.path.join(ROOT, "include", "rvn_gpu.h")]
    return {"librvn_gpu.so": _digest([os.path.join(csrc, f) for f in os.listdir(csrc)] + hdr, [os.path.basename(NVCC)] + NVCC_FLAGS),
            "libraven_b200.so": _digest([os.path.join(hdir, f) for f in os.listdir(hdir)] + hdr, ["g++", "-std=c++14", "-O2", "-fPIC", "-shared"])}


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("ok")
