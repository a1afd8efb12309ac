"""In-tree build recipes (explicit nvcc / g++ command lines; no JIT cache, no cmake).

Artefacts (all git-ignored, all travel to the GPU box with the gpurun snapshot):
  bang-billion-scale-ann_b200/libbang_b200.so      CUDA kernels + C ABI (sm_100a)         <- the product
  bang-billion-scale-ann_b200/libbang_b200_prof.so the same with per-phase clocks (BANG_B200_TIMERS=2 only)
  bang-billion-scale-ann_b200/libbang_fixture.so   host-only fixture builder (Vamana)     <- tooling
  bang-billion-scale-ann_b200/bang_search          CLI driver with the reference's argv   <- product
  oracle/libbang_oracle.so                         CPU restatement (test infrastructure)
  oracle/_ref/*                                    the unmodified reference, built from /root/reference
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
ORACLE = os.path.join(ROOT, "oracle")

LIB_CUDA = os.path.join(PKG_DIR, "libbang_b200.so")
LIB_FIXTURE = os.path.join(PKG_DIR, "libbang_fixture.so")
CLI = os.path.join(PKG_DIR, "bang_search")
CLI_INMEM = os.path.join(PKG_DIR, "bang")
LIB_PREPROCESS = os.path.join(PKG_DIR, "libbang_preprocess.so")
CLI_PREPROCESS = os.path.join(PKG_DIR, "bang_preprocess")
LIB_ORACLE = os.path.join(ORACLE, "libbang_oracle.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources if os.path.exists(s))


def _run(cmd: list[str], cwd: str | None = None) -> None:
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError(f"build step failed: {cmd[0]} (exit {r.returncode})")
    if os.environ.get("BANG_BUILD_VERBOSE"):
        sys.stderr.write(r.stdout)


def _srcs(d: str, exts=(".cu", ".cuh", ".cpp", ".h", ".c")) -> list[str]:
    out = []
    for base, _, files in os.walk(d):
        if "_ref" in base:
            continue
        out += [os.path.join(base, f) for f in files if f.endswith(exts)]
    return out


CUDA_SOURCES = ["bang_b200.cu", "builder.cu", "prep_kernels.cu", "filter_slots.cu", "search_inst_u8.cu", "search_inst_i8.cu", "search_inst_f32.cu"]
HOST_SOURCES = ["loader.cpp", "bang_shim.cpp", "shard_mem.cpp"]


def _compile_cuda_lib(out: str, defines: list[str], verbose_ptxas: bool) -> None:
    """nvcc -c per translation unit, in parallel (the search kernel is instantiated per element type in three of
    them), then one link step.  Objects go to a scratch directory under build/ (git-ignored)."""
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(PKG_DIR, "build", os.path.basename(out).replace(".so", ""))
    os.makedirs(objdir, exist_ok=True)
    common = [*ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fopenmp,-O3", *[f"-D{d}" for d in defines],
              "-I", INCLUDE, "-I", CSRC]
    if verbose_ptxas:
        common += ["-Xptxas", "-v"]
    objs = []
    jobs = []
    for src in CUDA_SOURCES + HOST_SOURCES:
        obj = os.path.join(objdir, src.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        jobs.append([NVCC, *common, "-c", os.path.join(CSRC, src), "-o", obj])
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
        list(ex.map(_run, jobs))
    # -Bsymbolic: references inside the library bind to its own definitions, so that libbang_b200_prof.so can be
    # dlopen'ed next to libbang_b200.so (same symbol names) and still launch its own kernels
    _run([NVCC, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-Xlinker", "-Bsymbolic", "-o", out, *objs, "-lgomp"])
    shutil.rmtree(objdir, ignore_errors=True)


def build_cuda(force: bool = False, verbose_ptxas: bool = False) -> str:
    deps = _srcs(CSRC) + _srcs(INCLUDE)
    if not force and _newer(LIB_CUDA, deps):
        return LIB_CUDA
    if not os.path.exists(NVCC):
        if os.path.exists(LIB_CUDA):
            return LIB_CUDA
        raise RuntimeError("nvcc not found and libbang_b200.so not prebuilt")
    _compile_cuda_lib(LIB_CUDA, [], verbose_ptxas)
    return LIB_CUDA


def build_variant(name: str, defines: list[str], verbose_ptxas: bool = False) -> str:
    """Debug / experimental builds of the CUDA library (not loaded by default): libbang_b200_<name>.so compiled with
    the given -D macros, e.g. build_variant("prof", ["BANG_PHASE_TIMERS"]).  Select at run time with
    BANG_B200_LIB=<path> (api.load_library)."""
    out = os.path.join(PKG_DIR, f"libbang_b200_{name}.so")
    _compile_cuda_lib(out, defines, verbose_ptxas)
    return out


LIB_PROF = os.path.join(PKG_DIR, "libbang_b200_prof.so")


def build_prof(force: bool = False) -> str:
    """libbang_b200_prof.so: the same sources with -DBANG_PHASE_TIMERS (per-phase SM clocks inside the fused kernel).
    The C++ class loads it instead of the product kernels when BANG_B200_TIMERS=2 and prints the reference's `_TIMERS`
    breakdown (bang_search.cu:1028-1051); never loaded otherwise."""
    deps = _srcs(CSRC) + _srcs(INCLUDE)
    if not force and _newer(LIB_PROF, deps):
        return LIB_PROF
    if not os.path.exists(NVCC):
        return LIB_PROF
    _compile_cuda_lib(LIB_PROF, ["BANG_PHASE_TIMERS"], False)
    return LIB_PROF


def build_cli(force: bool = False) -> str:
    src = os.path.join(CSRC, "bang_search_main.cpp")
    if not force and _newer(CLI, [src, LIB_CUDA] + _srcs(INCLUDE)):
        return CLI
    _run(["g++", "-O2", "-std=c++17", "-fopenmp", "-I", INCLUDE, src, "-o", CLI,
          "-L", PKG_DIR, "-lbang_b200", f"-Wl,-rpath,{PKG_DIR}", "-Wl,-rpath,$ORIGIN"])
    return CLI


def build_cli_inmem(force: bool = False) -> str:
    """`bang`: the 15-argument command line of the reference's Inmemory / Exactdistance forks (parANN.cu:79-93)."""
    src = os.path.join(CSRC, "bang_inmem_main.cpp")
    if not force and _newer(CLI_INMEM, [src, LIB_CUDA] + _srcs(INCLUDE)):
        return CLI_INMEM
    _run(["g++", "-O2", "-std=c++17", "-I", INCLUDE, src, "-o", CLI_INMEM,
          "-L", PKG_DIR, "-lbang_b200", f"-Wl,-rpath,{PKG_DIR}", "-Wl,-rpath,$ORIGIN"])
    return CLI_INMEM


def build_preprocess(force: bool = False) -> str:
    """DiskANN .index -> BANG .bin converter (host only): shared library for ctypes + the CLI with the reference
    script's argument list (BANG_Base/bang_preprocess.py)."""
    src = os.path.join(CSRC, "bang_preprocess.cpp")
    if not force and _newer(LIB_PREPROCESS, [src]) and _newer(CLI_PREPROCESS, [src]):
        return LIB_PREPROCESS
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", src, "-o", LIB_PREPROCESS])
    _run(["g++", "-O2", "-std=c++17", "-DBANG_PREPROCESS_MAIN", src, "-o", CLI_PREPROCESS])
    return LIB_PREPROCESS


def build_fixture(force: bool = False) -> str:
    src = os.path.join(CSRC, "fixture_builder.cpp")
    if not force and _newer(LIB_FIXTURE, [src]):
        return LIB_FIXTURE
    _run(["g++", "-O3", "-std=c++17", "-fopenmp", "-fPIC", "-shared", src, "-o", LIB_FIXTURE])
    return LIB_FIXTURE


def build_oracle(force: bool = False) -> str:
    srcs = [os.path.join(ORACLE, "bang_oracle.c"), os.path.join(ORACLE, "bang_oracle.h")]
    if not force and _newer(LIB_ORACLE, srcs):
        return LIB_ORACLE
    # -ffp-contract=off: every fused multiply-add in the oracle is an explicit fmaf(), as on the GPU.
    _run(["gcc", "-O2", "-std=c11", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared",
          srcs[0], "-o", LIB_ORACLE, "-lm"])
    return LIB_ORACLE


def build_reference(force: bool = False) -> str | None:
    """Compile the UNMODIFIED reference (BANG_Base) from /root/reference into oracle/_ref/ when it is
    mounted (build container only).  Sources are compiled where they lie; nothing is copied."""
    script = os.path.join(ORACLE, "build_ref.sh")
    out = os.path.join(ORACLE, "_ref", "libbang.so")
    if not os.path.isdir("/root/reference"):
        return out if os.path.exists(out) else None
    if not force and os.path.exists(out) and os.path.exists(os.path.join(ORACLE, "_ref", "ref_driver")):
        return out
    if not os.path.exists(script):
        return None
    _run(["bash", script])
    return out


def build_all(force: bool = False) -> None:
    build_fixture(force)
    build_preprocess(force)
    build_oracle(force)
    build_cuda(force)
    build_prof(force)
    build_cli(force)
    build_cli_inmem(force)
    build_reference(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", LIB_CUDA, LIB_FIXTURE, LIB_ORACLE, CLI)
