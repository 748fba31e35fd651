"""Builds kmernator_b200/libkmernator_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "kmn_api.cu")
DEPS = [SRC, os.path.join(HERE, "csrc", "kmn_kernels.cuh"), os.path.join(HERE, "csrc", "kmn_device.cuh"),
        os.path.join(os.path.dirname(HERE), "include", "kmernator_b200.h")]
OUT = os.path.join(HERE, "libkmernator_b200.so")


def _nccl_flags():
    """Prefer the NCCL that torch bundles (the one loaded in the bench/test processes); fall back to the system one."""
    try:
        import nvidia.nccl as n  # type: ignore
        root = os.path.dirname(n.__file__) if getattr(n, "__file__", None) else list(n.__path__)[0]
        inc, lib = os.path.join(root, "include"), os.path.join(root, "lib")
        if os.path.exists(os.path.join(inc, "nccl.h")) and os.path.exists(os.path.join(lib, "libnccl.so.2")):
            return ["-DKMN_WITH_NCCL", "-I" + inc, "-L" + lib, "-l:libnccl.so.2", "-Xlinker", "-rpath=" + lib]
    except Exception:
        pass
    if os.path.exists("/usr/include/nccl.h"):
        return ["-DKMN_WITH_NCCL", "-lnccl"]
    return []


def build(force=False, verbose=False, variant=None, extra=()):
    """variant/extra: developer A/B builds (libkmernator_b200.<variant>.so with extra -D flags), selected at load time
    with KMN_LIB_VARIANT=<variant>; the default library is the product."""
    out = OUT if not variant else OUT[:-3] + "." + variant + ".so"
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in DEPS):
        return out
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
           "-Xcompiler", "-fPIC", "-shared", "-ccbin", "/usr/bin/g++", "-o", out, SRC] + _nccl_flags() + list(extra)
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
