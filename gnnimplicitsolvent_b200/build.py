"""Builds libgnnis.so (hand-written sm_100a CUDA behind the C ABI of include/gnnis.h) in-tree with nvcc.

No torch headers are involved: the library is plain CUDA + a C ABI, loaded through ctypes (_lib.py).
`python -m gnnimplicitsolvent_b200.build` or __graft_entry__.build()."""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libgnnis.so")
SOURCES = ["api.cu", "pair_kernels.cu", "cell_list.cu", "edge_mlp.cu", "edge_mlp_tc.cu", "node_mlp.cu", "node_mlp_tc.cu", "dynamics.cu", "forcefield.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found")
    return nvcc


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "gnnis.h"))
    objdir = os.path.join(CSRC, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc, *ARCH, *FLAGS, "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append((s, cmd))
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        futs = {ex.submit(subprocess.run, cmd, capture_output=True, text=True): s for s, cmd in jobs}
        for f in cf.as_completed(futs):
            r = f.result()
            if verbose and r.stderr:
                print(r.stderr, file=sys.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {futs[f]}:\n{r.stdout}\n{r.stderr}")
    objs = [os.path.join(objdir, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(OUT, objs):
        r = subprocess.run([nvcc, *ARCH, "-shared", "-o", OUT, *objs, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


TORCH_OP_OUT = os.path.join(HERE, "libgnnis_torch.so")


def build_torch_op(force: bool = False) -> str:
    """Compiles csrc/torch_op.cpp (TORCH_LIBRARY shim over the C ABI) with g++ against the installed libtorch and
    links it to libgnnis.so (rpath $ORIGIN).  No CUDA compiler is involved: the shim only forwards pointers."""
    import torch
    from torch.utils import cpp_extension as ce

    lib = build_library()
    src = os.path.join(CSRC, "torch_op.cpp")
    hdr = os.path.join(os.path.dirname(HERE), "include", "gnnis.h")
    if not (force or _stale(TORCH_OP_OUT, [src, hdr, lib])):
        return TORCH_OP_OUT
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = [f"-I{d}" for d in ce.include_paths()] + [f"-I{os.path.join(cuda_home, 'include')}",
                                                   f"-I{os.path.join(os.path.dirname(HERE), 'include')}"]
    tlib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           *inc, src, "-o", TORCH_OP_OUT, f"-L{HERE}", "-l:libgnnis.so", f"-L{tlib}", "-lc10", "-lc10_cuda", "-ltorch_cpu",
           "-ltorch_cuda", "-ltorch", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed on torch_op.cpp:\n{r.stdout}\n{r.stderr}")
    return TORCH_OP_OUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_torch_op(force="--force" in sys.argv))
