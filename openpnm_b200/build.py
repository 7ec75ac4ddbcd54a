"""In-tree build of libb200pnm.so (sm_100a) with nvcc.  No JIT cache: the
shared object lands next to this file so it travels with the repository
snapshot to the GPU box."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libb200pnm.so")
SOURCES = ["api.cu", "assembly.cu", "solver.cu", "picard.cu", "dist.cu", "p2p.cu", "topology.cu", "transient.cu", "bicgstab.cu",
           "amg.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _nccl_include():
    for cand in ("/usr/include",):
        if os.path.exists(os.path.join(cand, "nccl.h")):
            return cand
    try:
        import nvidia.nccl as m  # torch's bundled headers
        return os.path.join(os.path.dirname(m.__file__), "include")
    except Exception:
        return None


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(ROOT, "include", "b200pnm.h"))
    return any(os.path.getmtime(d) > t for d in deps)


LIB_DBG = os.path.join(HERE, "libb200pnm_dbg.so")


def build_library(force=False, verbose=False, bounds=False):
    """bounds=True: the debug build with in-kernel index asserts (-DB200_BOUNDS), written to
    libb200pnm_dbg.so and loaded instead of the normal library when B200_LIB=dbg."""
    if bounds:
        return _build(LIB_DBG, os.path.join(HERE, "build", "dbg"), verbose, ["-DB200_BOUNDS"])
    if not force and not needs_build():
        return LIB
    return _build(LIB, os.path.join(HERE, "build"), verbose, [])


def _build(LIB, objdir, verbose, extra):
    nvcc = _nvcc()
    os.makedirs(objdir, exist_ok=True)
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    ni = _nccl_include()
    if ni and ni != "/usr/include":
        inc += ["-I", ni]
    flags = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"] + ARCH + extra
    if verbose:
        flags += ["-Xptxas", "-v"]
    procs, objs = [], []
    for src in SOURCES:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, "-c", os.path.join(CSRC, src), "-o", obj] + flags + inc
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}:\n{out}\n")
        elif verbose or out.strip():
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("build of libb200pnm.so failed")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH + ["-lcudart", "-ldl"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv,
                        bounds="--bounds" in sys.argv))
