"""tcgen05 / TMEM / bulk-copy building blocks (tests/probe_umma.cu) on a real B200."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_umma_probe():
    exe = os.path.join(ROOT, "build", "probe_umma")
    src = os.path.join(ROOT, "tests", "probe_umma.cu")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
                               "-std=c++17", "-o", exe, src])
    p = subprocess.run(["timeout", "120", exe], capture_output=True, text=True)
    print(p.stdout, p.stderr)
    assert "PROBE_SUMMARY ALL_PASS" in p.stdout, p.stdout + p.stderr


@pytest.mark.gpu
def test_tmem_operand_probe():
    """A operand in tensor memory (TS tcgen05.mma, written with tcgen05.st) == A in shared memory == CPU, bit for bit."""
    exe = os.path.join(ROOT, "build", "probe_ts")
    src = os.path.join(ROOT, "tests", "probe_ts.cu")
    deps = [src] + [os.path.join(ROOT, "codenerf_b200", "csrc", f) for f in ("umma.cuh", "sm100_common.cuh")]
    if not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
                               "-I", os.path.join(ROOT, "include"), "-o", exe, src, "-lcuda"])
    p = subprocess.run(["timeout", "120", exe, "--check-only"], capture_output=True, text=True)
    print(p.stdout, p.stderr)
    assert "TS_PROBE PASS" in p.stdout, p.stdout + p.stderr
