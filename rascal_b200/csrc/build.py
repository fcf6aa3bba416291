"""Build librascal_b200.so (sm_100a only) in-tree with nvcc.

    python -m rascal_b200.csrc.build

The shared library lands next to the package (rascal_b200/librascal_b200.so) so
it travels with the repo snapshot to the GPU box; nothing is JIT-compiled at
import time.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT = os.path.join(PKG, "librascal_b200.so")
SOURCES = ["abi.cu", "hough.cu", "vote.cu", "ransac.cu", "refit.cu", "batch.cu", "peaks.cu"]
HEADERS = ["common.cuh", os.path.join(PKG, "..", "include", "rascal_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--fmad=true",  # bit-exact paths use explicit __dmul_rn/__dadd_rn intrinsics
]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, s) for s in SOURCES] + [
        h if os.path.isabs(h) else os.path.join(HERE, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, variant=None, defines=()):
    """variant: build a second copy of the library with extra -D defines (tuning knobs such as
    RB2_A_CTAS) into rascal_b200/_variants/librascal_b200_<variant>.so; RB2_LIB=<path> selects it at load
    time (A/B timing on the GPU box, never the default)."""
    out = OUT
    if variant:
        vdir = os.path.join(PKG, "_variants")
        os.makedirs(vdir, exist_ok=True)
        out = os.path.join(vdir, "librascal_b200_%s.so" % variant)
        force = True
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    bdir = os.path.join(HERE, "_build" + ("_" + variant if variant else ""))
    os.makedirs(bdir, exist_ok=True)
    for s in SOURCES:
        o = os.path.join(bdir, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *["-D" + d for d in defines], *os.environ.get("RB2_NVCC_EXTRA", "").split(), "-c",
               os.path.join(HERE, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    for s, p in procs:
        log, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(log)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {s}")
    cmd = [nvcc, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    var = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var,
                defines=[a[2:] for a in sys.argv if a.startswith("-D")]))
