"""Build libcodenerf_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

`python -m codenerf_b200.build [--force] [-v] [--trace]`.  `--trace` builds the instrumented variant
(`-DCNB_TRACE`, clock64 accounting of the pipeline roles) as libcodenerf_b200_trace.so next to the product
library; scripts select it with `CNB_LIB=trace` (see _lib.py).  The product library never contains the
instrumentation.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcodenerf_b200.so")
LIB_TRACE = os.path.join(HERE, "libcodenerf_b200_trace.so")
SOURCES = ["api.cu", "rays.cu", "mlp_fp32.cu", "render_sm100.cu", "backward_sm100.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]


def _nvcc():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "codenerf_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, trace=False, variant=None, defines=()):
    """`variant` + `defines`: an experiment build libcodenerf_b200_<variant>.so compiled with the given -D flags
    (selected at run time with CNB_LIB=<variant>; A/B measurements in one GPU session)."""
    lib = LIB_TRACE if trace else LIB
    if variant:
        lib = os.path.join(HERE, f"libcodenerf_b200_{variant}.so")
    if not force and not needs_build(lib):
        return lib
    objs = []
    procs = []
    obj_dir = os.path.join(HERE, "..", "build", variant if variant else ("trace" if trace else ""))
    os.makedirs(obj_dir, exist_ok=True)
    extra = os.environ.get("CNB_NVCC_EXTRA", "").split() + (["-DCNB_TRACE"] if trace else []) + ["-D" + d for d in defines]
    for s in SOURCES:
        o = os.path.join(obj_dir, s.replace(".cu", ".o"))
        cmd = [_nvcc()] + extra + NVCC_FLAGS + ["-Xptxas", "-v" if verbose else "-O3", "-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(o)
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {s} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [_nvcc(), "-shared", "-o", lib] + objs + ["-lcudart"]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    var = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")]
    defs = [a[2:] for a in sys.argv if a.startswith("-D")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv,
                variant=var[0] if var else None, defines=defs))
