"""In-tree build of libr4d.so (hand-written CUDA for sm_100a) with plain nvcc.

The shared object lands next to this file (rag4dyg_b200/libr4d.so) so that it travels with the repo snapshot to the
GPU box and shows up in the process' loaded-library list.  nvcc cross-compiles without a GPU.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
OBJ_DIR = os.path.join(PKG_DIR, "build")
LIB_PATH = os.path.join(PKG_DIR, "libr4d.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libr4d.so cannot be built")
    return exe


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(PKG_DIR), "include", "r4d.h"))
    return max(os.path.getmtime(h) for h in hdrs)


def _compile_one(src, force):
    obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
    log = obj[:-2] + ".ptxas.log"
    if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(src)
            and os.path.getmtime(obj) >= _deps_mtime()):
        return obj, False
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    return obj, True


def build_lib(force=False, verbose=False):
    """Compile every csrc/*.cu for sm_100a and link rag4dyg_b200/libr4d.so.  Returns the library path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile_one(s, force), srcs))
    objs = [o for o, _ in results]
    rebuilt = any(ch for _, ch in results)
    if rebuilt or force or not os.path.exists(LIB_PATH):
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs + [
            "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        for o in objs:
            log = o[:-2] + ".ptxas.log"
            if os.path.exists(log):
                sys.stdout.write(open(log).read())
    return LIB_PATH


if __name__ == "__main__":
    p = build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", p)
