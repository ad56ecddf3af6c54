"""Build libacm_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with
the repo snapshot to the GPU box).  ``python -m acm_gnn_b200.build [--force]``"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libacm_b200.so")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["csr.cu", "gemm_simt.cu", "gemm.cu", "gemm_tc.cu", "fused_fwd.cu", "spmm_fwd.cu", "mix_bwd.cu", "spmm_t.cu", "loss.cu", "glue.cu", "params.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(os.path.dirname(HERE), "include", "acm_b200.h"), os.path.abspath(__file__)]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(src):
    obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
    srcp = os.path.join(CSRC, src)
    hdr_t = max(os.path.getmtime(d) for d in _deps() if d.endswith((".cuh", ".h", "build.py")))
    if os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(srcp), hdr_t):
        return obj, ""
    cmd = [NVCC] + FLAGS + ["-c", srcp, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    os.makedirs(OBJ_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        res = list(ex.map(_compile, SOURCES))
    log = "\n".join(r[1] for r in res)
    with open(os.path.join(OBJ_DIR, "ptxas.log"), "a" if not force else "w") as f:
        f.write(log)
    if verbose:
        print(log)
    cmd = [NVCC, "-shared", "-o", OUT] + [r[0] for r in res] + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
