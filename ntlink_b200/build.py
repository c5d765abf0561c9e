"""Build libntlink_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache: the .so travels with the repo)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libntlink_b200.so")
SOURCES = ["capi.cu", "sketch.cu", "map.cu", "scan.cu", "emit.cpp"]
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "ntlink_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "-Xptxas", "-v"]


def nvcc_path():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objdir = os.path.join(HERE, "_obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = nvcc_path()
    env = dict(os.environ)
    # /opt/gcc wrappers in this image lack some specs; use the system g++ as host compiler
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        cmd = [nvcc, "-ccbin", ccbin, "-c", os.path.join(CSRC, src), "-o", obj] + NVCC_FLAGS
        if src.endswith(".cpp"):
            cmd = [nvcc, "-ccbin", ccbin, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj] + NVCC_FLAGS
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src}\n{out}")
        if p.returncode != 0:
            failed = True
    text = "\n".join(log)
    with open(os.path.join(objdir, "build.log"), "w") as fout:
        fout.write(text)
    if failed:
        sys.stderr.write(text)
        raise RuntimeError("nvcc failed (see above)")
    cmd = [nvcc, "-ccbin", ccbin, "-shared", "-o", LIB] + objs + ["-lz", "-lpthread", "-gencode",
                                                                    "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd, env=env)
    if verbose:
        print(text)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
