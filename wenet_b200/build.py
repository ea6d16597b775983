"""Build libwenet_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m wenet_b200.build            # incremental
    python -m wenet_b200.build --force
The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# WB_BUILD_VARIANT=<name> builds a side-by-side variant (e.g. NVCC_EXTRA=-DWB_GEMM_DIAG WB_BUILD_VARIANT=diag) into
# lib/libwenet_b200_<name>.so; WENET_B200_LIB selects it at load time (tools only - the product loads the default)
_VAR = os.environ.get("WB_BUILD_VARIANT", "")
OBJ = os.path.join(HERE, "csrc", "_obj" + ("_" + _VAR if _VAR else ""))
LIB = os.path.join(HERE, "lib", "libwenet_b200%s.so" % ("_" + _VAR if _VAR else ""))

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    "-diag-suppress", "550",
] + os.environ.get("NVCC_EXTRA", "").split()


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_digest():
    h = hashlib.sha1()
    paths = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".h", ".cuh"))]
    paths.append(os.path.join(HERE, "..", "include", "wenet_b200.h"))
    for p in paths:
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    hd = _headers_digest()
    stamp = os.path.join(OBJ, "headers.sha1")
    old = open(stamp).read() if os.path.exists(stamp) else ""
    if old != hd:
        force = True
    todo, objs = [], []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < os.path.getmtime(s):
            todo.append((s, o))

    def _cc(so):
        s, o = so
        cmd = [NVCC] + FLAGS + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (s, r.stdout, r.stderr))
        return s

    if todo:
        if verbose:
            print("[wenet_b200.build] compiling %d file(s) for sm_100a" % len(todo), file=sys.stderr)
        with ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            list(ex.map(_cc, todo))
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp, "w") as fh:
        fh.write(hd)
    return LIB


def build_runtime(verbose: bool = True) -> str:
    """The C++ AsrModel back-end over the C ABI (runtime/b200_asr_model.cc) and its command-line driver."""
    root = os.path.join(HERE, "..", "runtime")
    lib = os.path.join(HERE, "lib", "libwenet_b200_runtime.so")
    exe = os.path.join(root, "b200_asr_main")
    src = [os.path.join(root, "b200_asr_model.cc"), os.path.join(root, "b200_asr_model.h"),
           os.path.join(HERE, "..", "include", "wenet_b200.h")]
    newest = max(os.path.getmtime(p) for p in src + [os.path.join(root, "b200_asr_main.cc"), LIB])
    if os.path.exists(lib) and os.path.exists(exe) and min(os.path.getmtime(lib), os.path.getmtime(exe)) >= newest:
        return exe
    cuda = os.path.dirname(os.path.dirname(NVCC))
    common = ["g++", "-O2", "-std=c++17", "-fPIC", "-I" + os.path.join(cuda, "include")]
    link = ["-L" + os.path.dirname(LIB), "-lwenet_b200", "-L" + os.path.join(cuda, "lib64"), "-lcudart",
            "-Wl,-rpath,$ORIGIN", "-Wl,-rpath,$ORIGIN/../wenet_b200/lib", "-Wl,-rpath," + os.path.join(cuda, "lib64")]
    cmds = [common + ["-shared", "-o", lib, src[0]] + link,
            common + ["-o", exe, os.path.join(root, "b200_asr_main.cc"), src[0]] + link]
    for cmd in cmds:
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("runtime build failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    if verbose:
        print("[wenet_b200.build] built %s and %s" % (lib, exe), file=sys.stderr)
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
    print(build_runtime())
