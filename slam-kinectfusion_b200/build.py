"""In-tree build of the native libraries (nvcc, sm_100a only)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def _needs(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _sources(d, exts):
    return [os.path.join(d, f) for f in sorted(os.listdir(d)) if f.endswith(exts)]


def build_all(force=False, verbose=False):
    """Compile libkfb200.so (CUDA kernels + C-ABI) and libkfusion_b200.so (C++ host facade)."""
    out = subprocess.DEVNULL if not verbose else None
    csrc = os.path.join(_HERE, "csrc")
    inc = os.path.join(os.path.dirname(_HERE), "include", "kfb200.h")
    lib = os.path.join(_HERE, "libkfb200.so")
    if force or _needs(lib, _sources(csrc, (".cu", ".cuh")) + [inc]):
        subprocess.check_call(["make", "-C", csrc] + (["-B"] if force else []), stdout=out)
    host = os.path.join(_HERE, "kfusion")
    hlib = os.path.join(_HERE, "libkfusion_b200.so")
    if os.path.isdir(host) and (force or _needs(hlib, _sources(os.path.join(host, "src"), (".cpp",)) +
                                                _sources(os.path.join(host, "include"), (".h", ".hpp")) + [inc, lib])):
        subprocess.check_call(["make", "-C", host] + (["-B"] if force else []), stdout=out)
    return lib
