"""TEST INFRASTRUCTURE: builds the CPU emulation of the CUDA kernels (g++ -DXW_EMU) used by the
`-m "not gpu"` kernel-logic tests.  Never used by the product package."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "xnode-wan-pde-solver_b200", "csrc")
OUT = os.path.join(HERE, "libxw_emu.so")


def build(force=False):
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(HERE, "cuda_emu.h"), os.path.join(ROOT, "include", "xnode_wan_b200.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(p) <= os.path.getmtime(OUT) for p in deps):
        return OUT
    cmd = ["g++", "-O2", "-std=c++20", "-DXW_EMU", "-x", "c++", "-I", HERE, "-I", CSRC, "-shared", "-fPIC",
           os.path.join(CSRC, "xw_capi.cu"), "-o", OUT, "-lpthread"]
    subprocess.run(cmd, check=True)
    return OUT
