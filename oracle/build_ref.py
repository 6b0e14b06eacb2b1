"""Compile the REFERENCE's own CUDA ops into oracle/_ref (test infrastructure only).

The reference (AIR-DISCOVER/Omni-PQ) ships its PointNet++ ops as a torch C++/CUDA extension
(`pointnet2/_ext_src`, pybind module `pointnet2._ext`, bindings.cpp:11-24).  This recipe compiles
those sources *where they lie* under /root/reference -- nothing is copied into the repo -- with
plain nvcc / g++ commands for sm_100a, and writes exactly one file:

    oracle/_ref/pn2_ref_ext.so        (git-ignored; travels to the GPU box with the snapshot)

It is used only by tests/ (reference-vs-oracle and reference-vs-ours parity on the B200 box) and by
bench.py's informational `ref_gpu` leg.  The product never loads it.  The reference's own build
system (setup.py) is not run.

Usage:  python oracle/build_ref.py [--force]
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/pointnet2/_ext_src"
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "pn2_ref_ext.so")
NAME = "pn2_ref_ext"


def available() -> bool:
    return os.path.isdir(REF_SRC)


def build(force: bool = False) -> str:
    if not available():
        raise FileNotFoundError(f"{REF_SRC} not present (only exists in the build container)")
    if os.path.exists(OUT_SO) and not force:
        return OUT_SO
    import torch
    from torch.utils import cpp_extension as ce

    os.makedirs(OUT_DIR, exist_ok=True)
    obj_dir = os.path.join(OUT_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    incs = [os.path.join(REF_SRC, "include")] + ce.include_paths() + [sysconfig.get_paths()["include"]]
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    incs.append(os.path.join(cuda_home, "include"))
    inc_flags = [f"-I{p}" for p in incs]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = [f"-DTORCH_EXTENSION_NAME={NAME}", "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={abi}"]
    srcs = sorted(os.listdir(os.path.join(REF_SRC, "src")))
    jobs = []
    for s in srcs:
        src = os.path.join(REF_SRC, "src", s)
        obj = os.path.join(obj_dir, s + ".o")
        if s.endswith(".cu"):
            # setup.py:25-28 of the reference passes only -O2; the arch is ours (sm_100a).
            cmd = [os.path.join(cuda_home, "bin", "nvcc"), "-O2", "-std=c++17", "-c", src, "-o", obj,
                   "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                   "-ccbin", "/usr/bin/g++",
                   "--expt-relaxed-constexpr", "-w"] + common + inc_flags
        elif s.endswith(".cpp"):
            cmd = ["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-w", "-c", src, "-o", obj] + common + inc_flags
        else:
            continue
        jobs.append((cmd, obj))

    def run(job):
        subprocess.run(job[0], check=True)
        return job[1]

    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        objs = list(ex.map(run, jobs))
    lib_dir = os.path.join(os.path.dirname(torch.__file__), "lib")
    link = ["/usr/bin/g++", "-shared", "-o", OUT_SO] + objs + [
        f"-L{lib_dir}", "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
        f"-L{os.path.join(cuda_home, 'lib64')}", "-lcudart", f"-Wl,-rpath,{lib_dir}"]
    subprocess.run(link, check=True)
    return OUT_SO


def load():
    """Import the compiled reference extension (needs torch imported first). None if not built."""
    if not os.path.exists(OUT_SO):
        return None
    import importlib.util
    import torch  # noqa: F401  (the .so links against libtorch)
    spec = importlib.util.spec_from_file_location(NAME, OUT_SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
