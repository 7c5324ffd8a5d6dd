"""Build libpvtrace_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m pvtrace_b200.csrc.build [--force] [--verbose]

The shared library travels to the GPU box with the repository snapshot; nothing is JIT-compiled at run time.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libpvtrace_b200.so")
SOURCES = ["pvt_api.cu"]
HEADERS = ["pvt_common.cuh", "pvt_rng.cuh", "pvt_math.cuh", "pvt_scene.cuh", "pvt_photon.cuh", "pvt_kernels.cuh",
           os.path.join(ROOT, "include", "pvtrace_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    for candidate in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if candidate and os.path.exists(candidate):
            return candidate
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale():
    if not os.path.exists(LIB):
        return True
    built = os.path.getmtime(LIB)
    deps = [os.path.join(HERE, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path(), *ARCH, "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
           "-Xcompiler", "-O2", "-o", LIB] + [os.path.join(HERE, s) for s in SOURCES]
    cmd[1:1] = os.environ.get("PVT_NVCC_FLAGS", "").split()  # e.g. -DPVT_PROFILE_STAGES (tools/stage_profile.sh)
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    # the image's default host compiler wrapper is fine for nvcc, but make sure a system gcc is visible
    env.setdefault("PATH", "/usr/bin:/bin")
    proc = subprocess.run(cmd, cwd=HERE, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
    if verbose:
        print(proc.stdout)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
