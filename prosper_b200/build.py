"""Build the CUDA library in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIBPATH = os.path.join(LIBDIR, "libprosper_b200.so")
SOURCES = ["engine.cu", "dgemm.cu", "gl_kernel.cu", "gl_state_tc.cu", "mca_kernel.cu", "gsc_kernel.cu", "solve.cu", "misc.cu", "ozaki.cu", "infer.cu", "synth.cu", "mixture.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]
OBJDIR = os.path.join(HERE, "build")


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force=False, verbose=False):
    """Compile prosper_b200/lib/libprosper_b200.so; returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and os.path.exists(LIBPATH) and os.path.getmtime(LIBPATH) >= _newest_source_mtime():
        return LIBPATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJDIR, exist_ok=True)
    hdr_m = max(os.path.getmtime(os.path.join(r, f)) for r in (CSRC, os.path.join(HERE, "..", "include"))
                for f in os.listdir(r) if f.endswith((".cuh", ".h", ".inc")))

    def compile_one(src):
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(hdr_m, os.path.getmtime(os.path.join(CSRC, src))):
            return obj
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, cwd=CSRC, check=True)
        return obj

    # one nvcc per translation unit, in parallel (objects under prosper_b200/build/ are git-ignored)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        objs = list(pool.map(compile_one, SOURCES))
    subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIBPATH] + objs, cwd=CSRC, check=True)
    return LIBPATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
