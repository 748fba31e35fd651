"""Builds the C++ host side (kmernator_b200/host/bin/FilterReads) against libkmernator_b200.so with g++."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
BIN = os.path.join(HERE, "bin")
FILTER_READS = os.path.join(BIN, "FilterReads")
FILTER_READS_P = os.path.join(BIN, "FilterReads-P")
MERACULOUS_COUNTER = os.path.join(BIN, "MeraculousCounter")


def _deps():
    d = [os.path.join(HERE, "apps", f) for f in ("FilterReads.cpp", "FilterReads-P.cpp", "FilterReads.h", "MeraculousCounter.cpp")]
    d.append(os.path.join(os.path.dirname(PKG), "include", "kmernator_b200.h"))
    kd = os.path.join(HERE, "kmernator")
    return d + [os.path.join(kd, f) for f in os.listdir(kd) if f.endswith(".h")]


def build(force=False):
    lib = os.path.join(PKG, "libkmernator_b200.so")
    if not os.path.exists(lib):
        raise ImportError("kmernator_b200.host: %s not built (build the CUDA library first)" % lib)
    os.makedirs(BIN, exist_ok=True)
    for exe, src in ((FILTER_READS, "FilterReads.cpp"), (FILTER_READS_P, "FilterReads-P.cpp"), (MERACULOUS_COUNTER, "MeraculousCounter.cpp")):
        if not force and os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in _deps()):
            continue
        cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-o", exe, os.path.join(HERE, "apps", src),
               "-L" + PKG, "-l:libkmernator_b200.so", "-Wl,-rpath,$ORIGIN/../..", "-Wl,-rpath-link," + PKG]
        subprocess.check_call(cmd)
    return FILTER_READS


if __name__ == "__main__":
    print(build(force=True))
