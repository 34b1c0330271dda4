"""Build the in-tree native artefacts of emme_b200 with nvcc for sm_100a.

  emme_b200/lib/libemme_b200.so   C ABI (include/emme_b200.h): CUDA kernels + host helpers
  emme_b200/bin/emme              C++ host program mirroring the reference's main()

nvcc cross-compiles without a GPU, so this runs in the CPU-only build container; the built
files are git-ignored but travel to the GPU box with the repository snapshot.
"""
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
LIB = PKG / "lib" / "libemme_b200.so"
EXE = PKG / "bin" / "emme"

CU_SOURCES = [PKG / "csrc" / n for n in ("assembly.cu", "dense.cu", "peer.cu", "qr.cu", "pic.cu", "capi.cu")]
HOST_SOURCES = [PKG / "host" / n for n in ("json.cpp", "parameters.cpp")]
EXE_SOURCES = [PKG / "host" / n for n in ("eigen_solver.cpp", "pic_solver.cpp", "main.cpp")]
HEADERS = (list((PKG / "csrc").glob("*.h")) + list((PKG / "csrc").glob("*.cuh")) +
           list((PKG / "host").glob("*.hpp")) + [ROOT / "include" / "emme_b200.h"])

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: emme_b200 needs the CUDA toolkit to build")
    return exe


def _stale(target, sources):
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources if Path(s).exists())


def _run(cmd, verbose):
    if verbose:
        print(" ".join(map(str, cmd)), flush=True)
    r = subprocess.run(list(map(str, cmd)), capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(map(str, cmd[:3])) + " ...")
    return r


def build_library(force=False, verbose=False, extra_flags=(), out=None, tag=""):
    srcs = CU_SOURCES + HOST_SOURCES
    LIB = out or globals()["LIB"]
    if not force and not _stale(LIB, srcs + HEADERS + [Path(__file__)]):
        return LIB
    LIB.parent.mkdir(exist_ok=True)
    objdir = PKG / "build" / tag if tag else PKG / "build"
    objdir.mkdir(exist_ok=True, parents=True)
    objs = []
    procs = []
    for src in srcs:
        obj = objdir / (src.stem + ".o")
        objs.append(obj)
        cmd = [_nvcc(), *NVCC_FLAGS, *extra_flags, "-I", ROOT / "include", "-c", src, "-o", obj]
        if verbose:
            print(" ".join(map(str, cmd)), flush=True)
        procs.append((cmd, subprocess.Popen(list(map(str, cmd)), stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("build failed: " + str(cmd[-3]))
    _run([_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"], verbose)
    return LIB


def build_executable(force=False, verbose=False):
    if not all(s.exists() for s in EXE_SOURCES):
        return None
    if not force and not _stale(EXE, EXE_SOURCES + HEADERS + [LIB]):
        return EXE
    EXE.parent.mkdir(exist_ok=True)
    _run(["/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++", "-O2", "-std=c++17",
          "-I", ROOT / "include", *EXE_SOURCES, PKG / "host" / "json.cpp", PKG / "host" / "parameters.cpp",
          "-o", EXE, "-L", LIB.parent, "-lemme_b200", f"-Wl,-rpath,$ORIGIN/../lib"], verbose)
    return EXE


def build_all(force=False, verbose=False):
    lib = build_library(force=force, verbose=verbose)
    exe = build_executable(force=force, verbose=verbose)
    return lib, exe


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))
