"""Builds srl_zoo_b200/csrc/libsrlz.so with nvcc for sm_100a (in-tree; the .so travels to the GPU box)."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libsrlz.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(os.path.dirname(HERE), "include", "srlz.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, dev=False):
    """dev=True: a separate libsrlz_dev.so compiled with -DSRLZ_DEV (clock64 timeline hooks for tools/dev_timeline.py);
    the product library has no debug entry points and no mode switches."""
    if dev:
        return _build_dev(verbose)
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    for src in sources():
        obj = src[:-3] + ".o"
        cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    ok = True
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("== %s\n%s\n" % (os.path.basename(src), out))
        ok &= p.returncode == 0
    if not ok:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return LIB


def _build_dev(verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    out = os.path.join(CSRC, os.environ.get("SRLZ_DEV_OUT", "libsrlz_dev.so"))     # variants for A/B runs: extra -D flags, own name
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + ["-DSRLZ_DEV"] + os.environ.get("SRLZ_DEV_DEFS", "").split()
    cmd = [nvcc] + flags + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", out] + sources() + ["-lcudart"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, dev="--dev" in sys.argv))
