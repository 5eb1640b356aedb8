"""In-tree build of the native libraries (nvcc for sm_100a, g++ for the host layer).

    python -m fastsmc_b200.build [--force]

Outputs go to fastsmc_b200/lib/ (git-ignored, but they travel with the gpurun snapshot).
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _host_cxx():
    # /opt/gcc/bin/g++ (the image's CXX) lacks some spec files; the distro compiler is complete.
    for c in ("/usr/bin/g++", shutil.which("g++") or ""):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("no g++ found")


def _nvcc():
    for c in (shutil.which("nvcc") or "", "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(*dirs, exts=(".cu", ".cuh", ".cpp", ".hpp", ".h")):
    out = []
    for d in dirs:
        for base, _, files in os.walk(d):
            out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return out


def build_device(force=False, verbose=False):
    """libfastsmc_b200.so: CUDA kernels + the C ABI of include/fastsmc_b200.h.  Every .cu under csrc/ is one
    translation unit; they compile concurrently and are linked into one shared library."""
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    target = os.path.join(LIBDIR, "libfastsmc_b200.so")
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = _sources(CSRC, os.path.join(ROOT, "include"))
    if force or _newer(target, deps):
        flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
        procs = []
        for src in srcs:
            obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
            procs.append((src, subprocess.Popen([_nvcc(), "-ccbin", _host_cxx()] + flags + ["-c", "-o", obj, src])))
        failed = [src for src, pr in procs if pr.wait() != 0]
        if failed:
            raise RuntimeError("nvcc failed on " + ", ".join(failed))
        objs = [os.path.join(objdir, os.path.basename(src)[:-3] + ".o") for src in srcs]
        subprocess.check_call([_nvcc(), "-ccbin", _host_cxx(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared",
                               "-o", target] + objs)
    return target


def build_host(force=False):
    """libfastsmc_host.so (C++ host classes over the C ABI) and the pyASMC extension module, if present."""
    hostdir = os.path.join(CSRC, "host")
    if not os.path.isdir(hostdir):
        return None
    os.makedirs(LIBDIR, exist_ok=True)
    cpps = sorted(f for f in _sources(hostdir, exts=(".cpp",)) if not f.endswith("pybind_module.cpp")
                  and not f.endswith("_main.cpp"))
    target = os.path.join(LIBDIR, "libfastsmc_host.so")
    deps = _sources(hostdir, os.path.join(ROOT, "include"))
    flags = ["-std=c++17", "-O2", "-fPIC", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", hostdir]
    if cpps and (force or _newer(target, deps)):
        subprocess.check_call([_host_cxx()] + flags + ["-shared", "-o", target] + cpps +
                              ["-L", LIBDIR, "-lfastsmc_b200", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"])
    # executables
    for main in sorted(f for f in _sources(hostdir, exts=(".cpp",)) if f.endswith("_main.cpp")):
        exe = os.path.join(LIBDIR, os.path.basename(main)[:-len("_main.cpp")] + "_exe")
        if force or _newer(exe, deps):
            subprocess.check_call([_host_cxx()] + flags + ["-o", exe, main, "-L", LIBDIR, "-lfastsmc_host",
                                                           "-lfastsmc_b200", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"])
    pyb = os.path.join(hostdir, "pybind_module.cpp")
    if os.path.exists(pyb):
        import pybind11
        ext = sysconfig.get_config_var("EXT_SUFFIX")
        mod = os.path.join(LIBDIR, "pyASMC" + ext)
        if force or _newer(mod, deps):
            subprocess.check_call([_host_cxx()] + flags + ["-shared", "-fvisibility=hidden", "-I", pybind11.get_include(),
                                                           "-I", sysconfig.get_paths()["include"], "-o", mod, pyb,
                                                           "-L", LIBDIR, "-lfastsmc_host", "-lfastsmc_b200",
                                                           "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread"])
    return target


def build_all(force=False, verbose=False):
    dev = build_device(force, verbose)
    host = build_host(force)
    return dev, host


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose="-v" in sys.argv))
