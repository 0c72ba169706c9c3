"""Build the in-tree native libraries of psdr_cuda_b200 (nvcc, sm_100a only).

  lib/libpsdr_b200.so   CUDA kernels + the C ABI of include/psdr_b200.h
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-Xptxas", "-v"]


def _stale(out, srcs):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(s) > t for s in srcs)


def build_core(force=False, verbose=False):
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libpsdr_b200.so")
    cu = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]
    cpp = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cpp")]
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "psdr_b200.h")]
    if not force and not _stale(out, deps):
        return out
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for src in cu + cpp:
        obj = os.path.join(HERE, "build", os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        o, _ = p.communicate()
        log.append(o)
        if p.returncode != 0:
            sys.stderr.write(o)
            raise RuntimeError("nvcc failed on %s" % src)
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + ["--cudart", "static"])
    return out


def build_host(force=False):
    """pybind11 host module psdr_cuda/_psdr_host*.so (g++), linked against lib/libpsdr_b200.so through an $ORIGIN rpath"""
    import sysconfig
    import pybind11
    core = build_core()
    out_dir = os.path.join(HERE, "compat", "psdr_cuda")
    out = os.path.join(out_dir, "_psdr_host" + sysconfig.get_config_var("EXT_SUFFIX"))
    srcs = [os.path.join(HERE, "host", f) for f in os.listdir(os.path.join(HERE, "host"))] + [os.path.join(ROOT, "include", "psdr_b200.h")]
    if not force and not _stale(out, srcs + [core]):
        return out
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-fvisibility=hidden",
           "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
           os.path.join(HERE, "host", "psdr_module.cpp"), "-o", out, "-L" + LIB, "-lpsdr_b200", "-Wl,-rpath,$ORIGIN/../../lib"]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_core(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host(force="--force" in sys.argv))
