"""Build the CUDA library IN-TREE for sm_100a.

    python assemble-them-all_b200/build.py [--force] [--verbose]

Output: assemble-them-all_b200/libsdfcontact_b200.so (git-ignored; it travels to
the GPU box with the gpurun snapshot). nvcc cross-compiles without a GPU.
The bit-critical part of the path is written with round-to-nearest intrinsics
that nvcc never contracts into FMAs (the reference's x86-64 build has none); the
remaining algebra may use FMAs. See csrc/sdfc_device.cuh.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("SDFC_OUT") or os.path.join(HERE, "libsdfcontact_b200.so")
SOURCES = ["sdfc_kernels.cu", "sdfc_capi.cu", "sdfc_step.cu", "sdfc_ground.cu", "sdfc_sdfgen.cu", "sdfc_microbench.cu"]
# sdfgen's fp32 / fp64 arithmetic has to match the reference's x86-64 build bit for bit: no FMA contraction
PER_SOURCE_FLAGS = {"sdfc_sdfgen.cu": ["-fmad=false"]}
FORCE_INST = "sdfc_force_inst.cu"  # compiled once per variant, -DSDFC_FORCE_VARIANT=0..7
HEADERS = ["sdfc_device.cuh", "sdfc_kernels.cuh", "sdfc_force.cuh", "sdfc_host.h", FORCE_INST,
           os.path.join("..", "..", "include", "sdfcontact_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--compiler-bindir", "/usr/bin/g++",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found")
    return nvcc


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if not force and not _stale():
        return OUT
    objdir = os.path.join(HERE, "build", os.path.basename(OUT))
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    min_blocks = os.environ.get("SDFC_FORCE_MIN_BLOCKS")  # tuning knob: __launch_bounds__(128, n) of k_force
    extra_all = [f"-DSDFC_FORCE_MIN_BLOCKS={int(min_blocks)}"] if min_blocks else []
    extra_all += os.environ.get("SDFC_EXTRA_FLAGS", "").split()  # tuning builds, e.g. "-DSDFC_LDG256=0"
    nf_blocks = os.environ.get("SDFC_FORCE_NF_MIN_BLOCKS")  # same knob for the planner fast path (k_force<1,0,0>)
    if nf_blocks:
        extra_all.append(f"-DSDFC_FORCE_NF_MIN_BLOCKS={int(nf_blocks)}")
    jobs = [(s, os.path.join(objdir, os.path.basename(s) + ".o"),
             list(extra_all) + PER_SOURCE_FLAGS.get(os.path.basename(s), [])) for s in srcs]
    jobs += [(os.path.join(CSRC, FORCE_INST), os.path.join(objdir, f"sdfc_force_v{v}.o"),
              [f"-DSDFC_FORCE_VARIANT={v}", *extra_all]) for v in range(8)]
    hdr_deps = [os.path.join(CSRC, h) for h in HEADERS if h != FORCE_INST] + [os.path.abspath(__file__)]
    flag_stamp = os.path.join(objdir, ".flags")
    flags_now = " ".join(NVCC_FLAGS + extra_all)
    flags_same = os.path.exists(flag_stamp) and open(flag_stamp).read() == flags_now
    for s, o, extra in jobs:
        objs.append(o)
        # per-object staleness: a .cu-only edit does not recompile the 8 force variants
        if not force and flags_same and os.path.exists(o) and all(
                os.path.getmtime(d) <= os.path.getmtime(o) for d in [s, *hdr_deps] if os.path.exists(d)):
            continue
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
    with open(flag_stamp, "w") as f:
        f.write(flags_now)
    cmd = [_nvcc(), "-shared", "-o", OUT, *objs, "--compiler-bindir", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
