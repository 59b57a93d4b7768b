"""Build libabides_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Two libraries come out of the same sources: the product (libabides_b200.so) and its -DABX_STRICT_SYNC twin (libabides_b200_strict.so: every
on-chip sync point a real __syncwarp()), which only the GPU parity suite loads -- every bit-exact test at production occupancy runs on both."""
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB = os.path.join(_PKG, "libabides_b200.so")
LIB_STRICT = os.path.join(_PKG, "libabides_b200_strict.so")
SOURCES = ["abx_sim.cu", "abx_qnet.cu"]
HEADERS = ["abx_core.cuh", "abx_warp.cuh", "abx_host_common.h", "abx_exp_table.h", os.path.join("..", "..", "include", "abides_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "-Xptxas", "-v"]


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def _start(lib, extra_flags):
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-o", lib] + [os.path.join(CSRC, s) for s in SOURCES]
    return cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def build(force=False, verbose=False, extra_flags=(), strict=True):
    jobs = []
    if force or needs_build(LIB):
        jobs.append((LIB, "build_ptxas.log") + _start(LIB, list(extra_flags)))
    if strict and (force or needs_build(LIB_STRICT)):
        jobs.append((LIB_STRICT, None) + _start(LIB_STRICT, ["-DABX_STRICT_SYNC", *extra_flags]))
    for lib, log, cmd, proc in jobs:                       # the two nvcc runs overlap
        out = proc.communicate()[0]
        if verbose or proc.returncode != 0:
            print(out)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed (%d): %s" % (proc.returncode, " ".join(cmd)))
        if log:
            with open(os.path.join(_PKG, log), "w") as f:
                f.write(out)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
