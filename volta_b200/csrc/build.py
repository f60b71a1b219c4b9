"""Builds libvolta_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m volta_b200.csrc.build [--force]
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
BUILD = os.path.join(HERE, "_build")
LIB = os.path.join(PKG, "libvolta_b200.so")
SOURCES = ["api.cu", "gemm_tc.cu", "gemm_tc_inst_a.cu", "gemm_tc_inst_b.cu", "gemm_tc_inst_c.cu", "gemm_tc_inst_d.cu",
           "gemm_tc_inst_e.cu", "gemm_simt.cu", "layernorm.cu", "attention.cu", "embed.cu", "loss.cu", "eltwise.cu", "optim.cu", "compact.cu", "pipeline.cu", "retrieval.cu", "allreduce.cu"]
HEADERS = ["common.cuh", "gemm_common.cuh", "gemm_tc_kernel.cuh", os.path.join("..", "..", "include", "volta_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: volta_b200 needs the CUDA 12.9 toolkit to build its sm_100a kernels")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    hdrs = [os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs = []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            jobs.append([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr)

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    objs = [os.path.join(BUILD, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lpthread", "-ldl"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
