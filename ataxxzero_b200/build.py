"""Builds csrc/*.cu into ataxxzero_b200/libataxxzero.so with plain nvcc for sm_100a.

In-tree on purpose: the .so is git-ignored but travels to the GPU box with the snapshot.
There is no torch extension and no JIT cache: the C ABI (include/ataxxzero.h) is the product
boundary, and Python binds it with ctypes (ataxxzero_b200/_native.py).
"""
import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libataxxzero.so")
OBJ = os.path.join(HERE, "build")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fno-strict-aliasing", "-Xptxas", "-v"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libataxxzero.so cannot be built")
    return exe


def host_compiler_args():
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); pin the distro compiler for nvcc
    return ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def _stamp(paths):
    h = hashlib.sha1()
    for p in sorted(paths):
        h.update(p.encode())
        h.update(open(p, "rb").read())
    h.update(" ".join(ARCH + NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) +
                     [os.path.join(HERE, "..", "include", "ataxxzero.h")])
    os.makedirs(OBJ, exist_ok=True)
    hdr_stamp = _stamp(headers)
    objs = []
    logs = []
    for src in sources:
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        stamp_file = obj + ".stamp"
        stamp = _stamp([src]) + hdr_stamp
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
            continue
        first = open(src).readline()
        extra = first.split("AZ_NVCC_FLAGS:", 1)[1].split() if "AZ_NVCC_FLAGS:" in first else []
        cmd = [nvcc()] + host_compiler_args() + ARCH + NVCC_FLAGS + extra + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs.append((src, r.stderr))
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed on %s" % src)
        open(stamp_file, "w").write(stamp)
        open(obj + ".ptxas.log", "w").write(r.stderr)
    need_link = force or not os.path.exists(OUT) or any(os.path.getmtime(o) > os.path.getmtime(OUT) for o in objs)
    if need_link:
        cmd = [nvcc()] + host_compiler_args() + ARCH + ["-shared", "-o", OUT] + objs
        if verbose:
            print(" ".join(cmd))
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
