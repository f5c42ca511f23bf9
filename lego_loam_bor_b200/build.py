"""Builds the CUDA shared library (C ABI of include/lego_loam_b200.h) for sm_100a with nvcc.

  python -m lego_loam_bor_b200.build [--force] [--verbose]

Flags: -gencode arch=compute_100a,code=sm_100a -lineinfo, and -fmad=false because bit-exact parity
with the reference's x86-64 (no FMA) arithmetic needs separately rounded multiplies and adds
(SURVEY.md section 10).  The .so is written in-tree (git-ignored) so that it travels with gpurun.
"""
import glob
import os
import subprocess
import sys

from ._paths import CSRC, INCLUDE, LIB_CUDA

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB_CUDA):
        return True
    t = os.path.getmtime(LIB_CUDA)
    from ._paths import PKG
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(INCLUDE, "*.h")) + glob.glob(os.path.join(PKG, "host", "*.cpp")) + \
        glob.glob(os.path.join(PKG, "host", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        from ._paths import PKG
        if not os.path.exists(os.path.join(PKG, "host", "sequence_driver")):
            build_host()
        from . import synth
        synth.build_cuda()
        return LIB_CUDA
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    logs = []
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(CSRC, "build", os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        # LEGO_LOAM_B200_NVCC_EXTRA: extra flags (e.g. -DSEG_TRIPS=8) for A/B builds
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("LEGO_LOAM_B200_NVCC_EXTRA", "").split() + ["-I", INCLUDE, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, pr in procs:
        out, _ = pr.communicate()
        logs.append(f"==== {os.path.basename(src)} ====\n{out}")
        if pr.returncode != 0:
            failed = True
    log_path = os.path.join(CSRC, "build", "ptxas.log")
    with open(log_path, "w") as f:
        f.write("\n".join(logs))
    if failed or verbose:
        sys.stderr.write("\n".join(logs))
    if failed:
        raise RuntimeError("nvcc failed; see " + log_path)
    subprocess.check_call([nvcc, "-shared", "-o", LIB_CUDA] + objs + ["-lcudart"])
    build_host()
    from . import synth
    synth.build_cuda()   # the device version of the synthetic-scan generator (bench / test data, a library of its own)
    return LIB_CUDA


def build_host():
    """The C++ host mirror of the reference's stage classes + the bag-less sequence driver."""
    from ._paths import PKG
    host = os.path.join(PKG, "host")
    exe = os.path.join(host, "sequence_driver")
    srcs = [os.path.join(host, "lego_loam_host.cpp"), os.path.join(host, "sequence_driver.cpp")]
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-g", "-pthread", "-I", INCLUDE] + srcs +
                          ["-L", PKG, "-llego_loam_b200", "-Wl,-rpath," + PKG, "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe])
    return exe


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(LIB_CUDA)
