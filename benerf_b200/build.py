"""Build libbenerf_b200.so in-tree with nvcc for sm_100a (no torch dependency in the library).

    python -m benerf_b200.build [--force]

The shared object is written to benerf_b200/lib/ (git-ignored, shipped to the GPU box by
gpurun).  pose.cu / rays.cu are compiled with -fmad=false: they sit upstream of the 2^9
positional-encoding frequency and follow the reference's separately rounded fp32 ops.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libbenerf_b200.so")
SOURCES = ["api.cu", "pose.cu", "rays.cu", "composite.cu", "image_formation.cu", "mlp_simt.cu", "mlp_tc.cu", "mlp_tc2.cu", "mlp_tc3.cu",
           "sgemm.cu", "gemm_tc.cu", "bwd_tiles.cu", "wgrad_pair.cu", "dgrad_chain.cu", "dgrad_chain2.cu", "backward.cu", "optim.cu", "probe_ts.cu", "loss.cu", "crf.cu"]
NO_FMAD = {"pose.cu", "rays.cu"}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "benerf_b200.h"))
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    nvcc = _nvcc()

    def compile_one(src):
        path = os.path.join(CSRC, src)
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        if not force and not _stale(obj, [path] + headers):
            return obj, ""
        cmd = [nvcc] + ARCH + COMMON + (["-fmad=false"] if src in NO_FMAD else []) + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=8) as ex:
        results = list(ex.map(compile_one, sources))
    objs = [o for o, _ in results]
    if verbose:
        for (_, log), src in zip(results, sources):
            if log:
                print(f"--- {src}\n{log}")
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB_PATH] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--force" in sys.argv))
