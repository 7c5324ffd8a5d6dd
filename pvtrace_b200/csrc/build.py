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
SOURCES = sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))  # one object per kernel family, compiled in parallel
HEADERS = sorted(f for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))) + [os.path.join(ROOT, "include", "pvtrace_b200.h")]
OBJDIR = os.path.join(HERE, "build")
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


def build(force=False, verbose=False, output=None, extra_flags=()):
    """Compile every .cu of this directory to an object (in parallel) and link them into the shared library."""
    import concurrent.futures

    lib = output or LIB
    if not force and output is None and not is_stale():
        return lib
    tag = os.path.splitext(os.path.basename(lib))[0]
    objdir = os.path.join(OBJDIR, tag)
    os.makedirs(objdir, exist_ok=True)
    env = dict(os.environ)
    env.setdefault("PATH", "/usr/bin:/bin")
    flags = [*ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
             *os.environ.get("PVT_NVCC_FLAGS", "").split(), *extra_flags]  # e.g. -DPVT_PROFILE_STAGES (tools/stage_profile.sh)
    if verbose:
        flags.append("-Xptxas=-v")

    def compile_one(src):
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc_path(), *flags, "-c", "-o", obj, os.path.join(HERE, src)]
        proc = subprocess.run(cmd, cwd=HERE, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
        return obj, proc.stdout

    with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            print(log)
    cmd = [nvcc_path(), *ARCH, "-shared", "-o", lib] + [obj for obj, _ in results]
    proc = subprocess.run(cmd, cwd=HERE, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("link failed:\n" + " ".join(cmd) + "\n" + proc.stdout)
    return lib


if __name__ == "__main__":
    out = None
    if "--output" in sys.argv:  # A/B variants: python -m pvtrace_b200.csrc.build --output lib_x.so -- -DFLAG ...
        out = os.path.join(HERE, sys.argv[sys.argv.index("--output") + 1])
    extra = sys.argv[sys.argv.index("--") + 1:] if "--" in sys.argv else []
    path = build(force="--force" in sys.argv or out is not None, verbose="--verbose" in sys.argv, output=out,
                 extra_flags=extra)
    print(path)
