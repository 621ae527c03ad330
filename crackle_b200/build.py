"""In-tree build of libcrackle_b200.so (CUDA kernels + C-ABI) and the pybind11 `fastcrackle` module.

nvcc cross-compiles for sm_100a without a GPU.  Objects are cached by mtime under crackle_b200/csrc/build/."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libcrackle_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CU_SOURCES = ["ckl_planes.cu", "ckl_trace.cu", "ckl_markov.cu", "ckl_labels.cu", "ckl_decode.cu", "ckl_api.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "crackle_b200.h"))
    return hs


def build_lib(verbose=False, force=False):
    os.makedirs(BUILD, exist_ok=True)
    hdrs = _headers()
    objs = []
    procs = []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src[:-3] + ".o")
        objs.append(o)
        if force or not _newer(o, [s] + hdrs):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if force or procs or not _newer(LIB, objs):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        subprocess.check_call(cmd)
    return LIB


def build_pymodule(force=False):
    """pybind11 `fastcrackle` drop-in (compress / decompress) above the C-ABI."""
    import pybind11
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(HERE, "fastcrackle" + ext)
    src = os.path.join(CSRC, "fastcrackle_module.cpp")
    if not force and _newer(target, [src, LIB] + _headers()):
        return target
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-I" + pybind11.get_include(),
           "-I" + sysconfig.get_paths()["include"], src, "-o", target, "-L" + HERE, "-lcrackle_b200",
           "-Wl,-rpath,$ORIGIN"]
    subprocess.check_call(cmd)
    return target


if __name__ == "__main__":
    build_lib(verbose="-v" in sys.argv, force="-f" in sys.argv)
    if os.path.exists(os.path.join(CSRC, "fastcrackle_module.cpp")):
        build_pymodule(force="-f" in sys.argv)
    print("built", LIB)
