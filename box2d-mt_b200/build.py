#!/usr/bin/env python3
"""Build libb2cuda.so (the sm_100a device library behind include/b2cuda.h) in-tree with nvcc.

    python box2d-mt_b200/build.py [--force] [--verbose]

-fmad=false: the parity contract needs every fp32 result to equal what the reference's x86-64 build
computes (no fused multiply-add there); -prec-div / -prec-sqrt keep IEEE division and sqrt.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb2cuda.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--shared", "-cudart", "shared",
]
SOURCES = ["prims.cu", "world.cu"]


def lib_path():
    return OUT


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "b2cuda.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    extra = os.environ.get("B2CU_NVCC_FLAGS", "").split()
    cmd = [NVCC] + FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", OUT]
    subprocess.run(cmd, check=True)
    return OUT


HOST = os.path.join(HERE, "host")
HOST_OUT = os.path.join(HERE, "libbox2d_b200.so")
HOST_SOURCES = ["src/b2Common.cpp", "src/b2Shapes.cpp", "src/b2Body.cpp", "src/b2World.cpp", "src/b2Joints.cpp", "src/b2Dump.cpp",
                "src/b2CudaStepExecutor.cpp", "src/b2CudaShardedWorld.cpp", "capi/b2host_capi.cpp"]


def host_lib_path():
    return HOST_OUT


def build_host(force=False):
    """libbox2d_b200.so: the C++ host API (b2World, b2Body, b2Fixture, b2CudaStepExecutor) + its flat C binding.
    -ffp-contract=off: host-computed mass data, AABBs and sincos must round like the reference's build."""
    srcs = [os.path.join(HOST, s) for s in HOST_SOURCES]
    deps = list(srcs) + [os.path.join(HERE, "..", "include", "b2cuda.h"), OUT]
    for d, _, files in os.walk(os.path.join(HOST, "Box2D")):
        deps += [os.path.join(d, f) for f in files]
    if not force and os.path.exists(HOST_OUT) and all(os.path.getmtime(d) <= os.path.getmtime(HOST_OUT) for d in deps):
        return HOST_OUT
    cmd = ["g++", "-std=c++11", "-O2", "-DNDEBUG", "-fPIC", "-shared", "-pthread", "-ffp-contract=off",
           "-Wall", "-I" + HOST, "-I" + os.path.join(HERE, "..", "include")] + srcs + \
          ["-L" + HERE, "-lb2cuda", "-Wl,-rpath,$ORIGIN", "-o", HOST_OUT]
    subprocess.run(cmd, check=True)
    return HOST_OUT


def build_all(force=False, verbose=False):
    build(force, verbose)
    build_host(force)
    return OUT, HOST_OUT


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
