"""In-tree build of libpolymlp_b200.so (CUDA kernels + C ABI) for sm_100a.

    python -m pypolymlp_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is written to pypolymlp_b200/lib/ (git-ignored,
shipped to the GPU box by gpurun).
"""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libpolymlp_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["pm_tables.cpp", "pm_kernels.cu", "pm_kernels_mma.cu", "pm_kernels_front.cu", "pm_kernels_feat.cu", "pm_solver.cu", "pm_capi.cu"]
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off",
    "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(HERE, "..", "include", "polymlp_b200.h"))
    return files


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(f) > t for f in _deps())


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.rsplit(".", 1)[0] + ".o")
        cmd = [NVCC] + FLAGS + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as f:
            f.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcublas", "-lcusolver", "-lcudart", "-ldl",
                                                 "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    build_pybind_module()
    return LIB


def pybind_module_path():
    import sysconfig

    return os.path.join(LIBDIR, "libmlpcpp" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_pybind_module():
    """pybind11 module `libmlpcpp` (host C++ above the C ABI), the drop-in for pypolymlp.cxx.lib.libmlpcpp."""
    import sysconfig

    import pybind11

    out = pybind_module_path()
    cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden",
           "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
           os.path.join(CSRC, "pybind_module.cpp"), "-L", LIBDIR, "-lpolymlp_b200",
           "-Wl,-rpath,$ORIGIN", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"pybind11 module build failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
