"""Build libplume_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python zk-nullifier-sig_b200/plume_b200/build.py [--force]

One object per kernel family, compiled in parallel, linked with a static CUDA runtime so the
library has no dependency beyond libstdc++ (it loads next to torch's own runtime without clashing).
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # zk-nullifier-sig_b200/
if PKG not in sys.path:
    sys.path.insert(0, PKG)
CSRC = os.path.join(PKG, "csrc")
BUILD = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libplume_b200.so")
UNITS = ["api", "api_multi", "selftest", "k_sign", "k_verify", "k_misc", "k_team", "k_team_lad"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-no-compress"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    d.append(os.path.join(os.path.dirname(PKG), "include", "plume_b200.h"))
    return d


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build(force=False, verbose=True, variant=None, defines=()):
    """variant/defines: build an experiment library libplume_b200_<variant>.so with extra -D flags
    (selected at run time with PLUME_B200_LIB=<path>); the default build is the shipped library."""
    global BUILD, LIB
    if variant:
        BUILD = os.path.join(PKG, "build", variant)
        LIB = os.path.join(PKG, "libplume_b200_%s.so" % variant)
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    deps = _deps()
    flags = NVCC_FLAGS + ["-D" + d for d in defines]

    def compile_unit(u):
        src, obj = os.path.join(CSRC, u + ".cu"), os.path.join(BUILD, u + ".o")
        if not force and not _stale(obj, [src] + deps):
            return u, False
        log = os.path.join(BUILD, u + ".log")
        with open(log, "w") as lf:
            r = subprocess.run([nvcc] + flags + ["-c", src, "-o", obj], stdout=lf, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (u, open(log).read()[-4000:]))
        return u, True

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        results = list(ex.map(compile_unit, UNITS))
    objs = [os.path.join(BUILD, u + ".o") for u in UNITS]
    if force or any(ch for _, ch in results) or _stale(LIB, objs):
        r = subprocess.run([nvcc] + ARCH + ["-shared", "-cudart", "static", "-o", LIB] + objs + ["-ldl", "-lpthread"],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout)
    if verbose:
        print("built", LIB, [u for u, ch in results if ch] or "(up to date)")
    return LIB


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    if args:   # build.py <variant> DEF1 DEF2=val ...
        build(force="--force" in sys.argv, variant=args[0], defines=args[1:])
    else:
        build(force="--force" in sys.argv)
