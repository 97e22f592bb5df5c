"""Build recipe for libimk.so (hand-written CUDA for sm_100a behind the C ABI of include/imk.h).

    python -m inconsistencymasks_b200.build [--force]

nvcc cross-compiles without a GPU.  The library is built IN-TREE
(inconsistencymasks_b200/libimk.so, git-ignored) so that it travels to the GPU box with
the repository snapshot.  cudart is linked statically: the .so has no dependency on the
CUDA runtime PyTorch bundles.
"""
from __future__ import annotations

import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("IMK_LIB_OUT") or os.path.join(HERE, "libimk.so")      # IMK_LIB_OUT + IMK_BUILD_FLAGS: instrumented variants
SOURCES = ["imk_api.cu", "imk_im.cu", "imk_morph.cu", "imk_unet.cu", "imk_conv_tc.cu", "imk_block_tc.cu", "imk_augment.cu", "imk_eval.cu", "imk_evalnet.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"] + os.environ.get("IMK_BUILD_FLAGS", "").split()
# objects of a variant build (extra flags) live in their own directory, so switching back does not recompile everything
OBJ = os.path.join(HERE, "csrc", "_obj" + ("_" + hashlib.sha256(os.environ["IMK_BUILD_FLAGS"].encode()).hexdigest()[:8]
                                            if os.environ.get("IMK_BUILD_FLAGS") else ""))


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libimk needs the CUDA toolkit (there is no CPU build)")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths, key=os.path.basename):
        h.update(os.path.basename(p).encode())      # names, not absolute paths: the digest is the same wherever the tree is mounted
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def sources_digest() -> str:
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "imk.h"))
    return _digest(deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link libimk.so.  Returns its path."""
    stamp = LIB + ".stamp"
    digest = sources_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)

    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    headers.append(os.path.join(os.path.dirname(HERE), "include", "imk.h"))

    def compile_one(src):
        # per-object stamp (source + every header + flags): an edit of one .cu recompiles that file only
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        want = _digest([os.path.join(CSRC, src)] + headers)
        ostamp = obj + ".stamp"
        if not force and os.path.exists(obj) and os.path.exists(ostamp) and open(ostamp).read().strip() == want:
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(ostamp, "w") as f:
            f.write(want)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
