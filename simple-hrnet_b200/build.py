"""In-tree build of libhrnet_b200.so (nvcc, sm_100a only).  Used by __graft_entry__.build() and by
the package on first import when the library is missing.  The .so is git-ignored but travels to the
GPU box with the gpurun snapshot."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhrnet_b200.so")
SOURCES = ["plan.cu", "conv_igemm_tc.cu", "conv3x3_patch_tc.cu", "conv_group.cu", "conv_chain.cu", "conv_xunit.cu", "stem_tc.cu", "simt_kernels.cu"]
HEADERS = ["hrnet_internal.h", "ptx.cuh", "epilogue.cuh", "conv3x3_patch_body.cuh", "conv_igemm_body.cuh", "chain_common.cuh", os.path.join("..", "..", "include", "hrnet_b200.h")]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compiles and links in-tree.  Several processes may get here at once (one rank per GPU under torchrun): an
    exclusive file lock serialises them and the losers find the library already up to date."""
    if not force and not needs_build():
        return LIB
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB
            return _build_locked(verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(verbose):
    objs = []
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((src, subprocess.Popen([_nvcc()] + flags + ["-c", os.path.join(CSRC, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    tmp = LIB + ".tmp"
    subprocess.check_call([_nvcc(), "-shared", "-Wno-deprecated-gpu-targets", "-o", tmp] + objs + ["-cudart", "static"])
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
