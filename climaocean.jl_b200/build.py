"""Build lib/libcoflux.so (CUDA, sm_100a) and oracle/libcoflux_oracle.so (C, test infrastructure)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SRC = os.path.join(HERE, "csrc", "coflux_abi.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "coflux_kernels.cuh"), os.path.join(HERE, "csrc", "coflux_solve_tile.cuh"), os.path.join(HERE, "csrc", "coflux_solve_stream.cuh"), os.path.join(HERE, "csrc", "coflux_psi_table.h"), os.path.join(HERE, "csrc", "coflux_fastmath.cuh"), os.path.join(HERE, "csrc", "coflux_math_tables.h"), os.path.join(HERE, "csrc", "coflux_device.cuh"),
        os.path.join(ROOT, "include", "coflux.h")]
OUT = os.path.join(HERE, "lib", "libcoflux.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-shared", "-diag-suppress", "177"]


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build_cuda(force=False, verbose=False, out=OUT, defines=()):
    """nvcc → lib/libcoflux.so.  `out`/`defines` build tuning variants (tools/ab_variants.py)."""
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if force or _stale(out, DEPS):
        cmd = [NVCC] + NVCC_FLAGS + [f"-D{d}" for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out, SRC]
        subprocess.run(cmd, check=True)
    return out


def build_tools(force=False):
    """lib/fm_check: device-side accuracy check of csrc/coflux_fastmath.cuh (run by tests/test_fastmath.py on the GPU box)."""
    src = os.path.join(ROOT, "tools", "fm_check.cu")
    out = os.path.join(HERE, "lib", "fm_check")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if force or _stale(out, [src, os.path.join(HERE, "csrc", "coflux_fastmath.cuh"), os.path.join(HERE, "csrc", "coflux_math_tables.h")]):
        subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-fmad=false", "-o", out, src], check=True)
    return out


def build_oracle(force=False):
    odir = os.path.join(ROOT, "oracle")
    out = os.path.join(odir, "libcoflux_oracle.so")
    deps = [os.path.join(odir, "coflux_oracle.c"), os.path.join(odir, "oracle_impl.h"), os.path.join(ROOT, "include", "coflux.h")]
    if force or _stale(out, deps):
        subprocess.run(["make", "-C", odir, "-B"], check=True, stdout=subprocess.DEVNULL)
    return out


if __name__ == "__main__":
    print(build_cuda(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_tools(force="--force" in sys.argv))
    print(build_oracle(force="--force" in sys.argv))
