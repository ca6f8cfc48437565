"""Build the test infrastructure: the C oracle (oracle/liboracle.so) and, when the reference tree is present, the
unmodified reference (oracle/_ref/, via build_ref.sh).  Building the checker is not using it: nothing under
ravenhydroframework_b200/ imports, links or executes anything from this directory."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def build(force=False):
    out = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "raven_oracle.c")
    hdr = os.path.join(ROOT, "include", "rvn_gpu.h")
    if force or not os.path.exists(out) or max(os.path.getmtime(src), os.path.getmtime(hdr)) > os.path.getmtime(out):
        subprocess.check_call(["gcc", "-std=c99", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), src, "-o", out, "-lm"])
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["bash", os.path.join(HERE, "build_ref.sh")])
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    print("ok")
