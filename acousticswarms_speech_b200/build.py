"""In-tree build of libasw.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m acousticswarms_speech_b200.build [--force] [--verbose]

The library is written to acousticswarms_speech_b200/lib/libasw.so so it travels with the tree.
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libasw.so")
STAMP = os.path.join(LIBDIR, "libasw.stamp")
SOURCES = ["api.cu", "stft_cc.cu", "stft_cc_warp.cu", "gcc.cu", "srp_gather.cu", "topk.cu", "shift_stack.cu", "prune.cu", "select.cu", "geometry.cu", "powers.cu", "xcorr.cu"]
# The warp-FFT kernel needs <= 128 registers/thread so that two CTAs fit the per-partition register files.
PER_FILE_FLAGS = {"stft_cc_warp.cu": ["-maxrregcount=128"]}
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; libasw.so cannot be built (there is no CPU fallback)")


def _digest():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))]
    files.append(os.path.join(HERE, "..", "include", "asw.h"))
    for f in files:
        if os.path.isfile(f):
            with open(f, "rb") as fh:
                h.update(f.encode())
                h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(repr(sorted(PER_FILE_FLAGS.items())).encode())
    return h.hexdigest()


def sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def build_lib(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == dig:
                return LIB
    nvcc = _nvcc()
    flags = list(NVCC_FLAGS)
    if verbose:
        flags += ["-Xptxas", "-v"]
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *flags, *PER_FILE_FLAGS.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed (see output above)")
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
