"""Build libvqb200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libvqb200.so")
SOURCES = ["abi.cu", "assign.cu", "assign_simt.cu", "assign_tc.cu", "assign_f16.cu", "assign_tc_gen.cu", "ema.cu", "peer.cu", "gather.cu", "tile_ops.cu", "rvq_small.cu", "fsq_lfq.cu", "fsq_lfq_fused.cu", "tokens.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-O2"]


def _nvcc():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def have_nvcc():
    try:
        _nvcc()
        return True
    except RuntimeError:
        return False


def build_locked(force=False):
    """build() under an exclusive file lock: under torchrun every rank may find the library missing or stale at once."""
    import fcntl
    os.makedirs(LIB_DIR, exist_ok=True)
    with open(os.path.join(LIB_DIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return build(force=force)            # whoever comes second finds an up-to-date library and returns at once
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


STAMP = os.path.join(LIB_DIR, "sources.sha256")


def source_hash():
    """Content hash of everything the library is built from (mtimes do not survive a copy of the tree)."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    deps += [os.path.join(HERE, "..", "include", "vqb200.h"), os.path.abspath(__file__)]
    for d in deps:
        if os.path.exists(d):
            h.update(os.path.basename(d).encode())
            with open(d, "rb") as f:
                h.update(f.read())
    h.update(os.environ.get("VQB200_NVCC_EXTRA", "").encode())
    return h.hexdigest()


def needs_build():
    """True when the library is missing or was linked from different sources.  A library WITHOUT a stamp (copied in
    from elsewhere) is taken as it is: rebuilding on a guess would cost minutes on every rank of a GPU job."""
    if not os.path.exists(LIB):
        return True
    if not os.path.exists(STAMP):
        return False
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link lib/libvqb200.so.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    flags = list(NVCC_FLAGS)
    flags += os.environ.get("VQB200_NVCC_EXTRA", "").split()          # development: e.g. -DVQB200_TOP2_VARIANT=0
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    procs = []
    for s in srcs:
        o = os.path.join(obj_dir, s.replace(".cu", ".o"))
        cmd = [nvcc] + flags + ["-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((s, o, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    for s, o, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose and out.strip():
            print(out)
        objs.append(o)
    tmp = LIB + f".tmp{os.getpid()}"
    cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB)                         # readers never see a partial library
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
