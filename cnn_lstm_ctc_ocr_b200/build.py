"""Build recipe for libocr_b200.so (hand-written sm_100a CUDA behind a C ABI).

`python -m cnn_lstm_ctc_ocr_b200.build` compiles every csrc/*.cu with
`nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo` and links them into
cnn_lstm_ctc_ocr_b200/libocr_b200.so IN-TREE (it is git-ignored but travels to the GPU box).
The CUDA runtime is linked statically so the library loads on a machine without a driver
(the CPU-only test tier checks the exported symbols); the driver API is reached through
cudaGetDriverEntryPoint at run time.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
SO = os.path.join(HERE, "libocr_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v", "--expt-relaxed-constexpr", "-Wno-deprecated-gpu-targets"]


def _nvcc():
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found; libocr_b200.so cannot be built")
    return nv


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdr.append(os.path.join(HERE, "..", "include", "ocr_b200.h"))
    hdr.append(os.path.abspath(__file__))
    return hdr


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    newest = max(os.path.getmtime(p) for p in [src] + _deps())
    if os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj, ""
    cmd = [_nvcc()] + ARCH + NVCC_FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = sources()
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    if verbose and log:
        sys.stderr.write(log)
    if log:
        with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
            f.write(log)
    if (not os.path.exists(SO)) or any(os.path.getmtime(o) > os.path.getmtime(SO) for o in objs):
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", SO] + objs + ["-cudart", "static", "-Xlinker", "--no-undefined"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return SO


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
