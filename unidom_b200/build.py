"""Build libunidom_b200.so in-tree with nvcc for sm_100a (no torch types, pure C ABI).

Usage: python -m unidom_b200.build [--force]
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libunidom_b200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# (source, extra flags).  The grid/cell units carry fp32 finite-difference normals: no FMA contraction.
UNITS = [
    ("abi.cu", []),
    ("mpm_particles.cu", []),
    ("mpm_grid.cu", ["--fmad=false"]),
    ("cloth.cu", []),
    ("reward.cu", []),
    ("apg_fused.cu", []),
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    for root, _, files in sorted(os.walk(CSRC)):
        for f in sorted(files):
            if f.endswith((".cu", ".cuh", ".h")):
                h.update(f.encode())
                h.update(open(os.path.join(root, f), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "unidom_b200.h"), "rb").read())
    h.update(repr((ARCH, COMMON, UNITS)).encode())
    return h.hexdigest()


def build(force=False, verbose=False, variant=None, defines=()):
    """variant/defines: development A/B builds, e.g. build(variant="hestenes", defines=["-DUD_SVD_HESTENES"]) writes
    libunidom_b200_hestenes.so, which `UNIDOM_B200_LIB=<path>` makes _lib.py load instead of the product library."""
    if variant:
        return _build(os.path.join(HERE, f"libunidom_b200_{variant}.so"), "." + variant, list(defines), verbose)
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read() == digest:
        return LIB
    _build(LIB, "", [], verbose)
    open(STAMP, "w").write(digest)
    return LIB


def _build(LIB, tag, defines, verbose):
    nvcc = _nvcc()
    objs = []
    procs = []
    for src, extra in UNITS:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(CSRC, src.replace(".cu", tag + ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, *defines, "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode:
            print(" ".join(cmd))
            print(out)
        if pr.returncode:
            raise RuntimeError("nvcc failed for " + cmd[-3])
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
    subprocess.check_call(cmd)
    return LIB


def build_xla_adapter(verbose=False):
    """csrc/xla_ffi.cc -> libunidom_b200_xla.so (the jax.ffi handlers) against jax's own XLA headers when jax is
    importable; otherwise -- this image has no jax -- against the test-only API stand-in under tests/xla_stub, as a
    compile check and for tests/test_xla_ffi.py (output next to the stub, never loaded by the product).
    Returns (path, "jax" | "stub")."""
    build()
    src = os.path.join(CSRC, "xla_ffi.cc")
    try:
        import jax.ffi
        inc, kind, out = jax.ffi.include_dir(), "jax", os.path.join(HERE, "libunidom_b200_xla.so")
    except Exception:
        stub = os.path.join(HERE, "..", "tests", "xla_stub")
        inc, kind, out = stub, "stub", os.path.join(stub, "libunidom_b200_xla_stub.so")
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(_nvcc_path()))), "include")
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-I", inc, "-I", cuda_inc, "-o", out, src,
           "-L", HERE, "-lunidom_b200", "-Wl,-rpath," + HERE]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out, kind


def _nvcc_path():
    import shutil
    n = _nvcc()
    return n if os.path.isabs(n) else (shutil.which(n) or "/usr/local/cuda/bin/nvcc")


if __name__ == "__main__":
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var,
                defines=[a for a in sys.argv[1:] if a.startswith("-D")]))
    if "--xla" in sys.argv:
        print(*build_xla_adapter(verbose="-v" in sys.argv))
