#!/usr/bin/env python
"""Build the UNMODIFIED reference (torchlsq 2.1) into oracle/_ref/ -- TEST INFRASTRUCTURE ONLY.

What this does
--------------
Compiles the reference's own C++/CUDA sources *where they lie* under /root/reference with
plain g++ / nvcc command lines (this script is the recipe; the reference's setup.py is NOT
run), and writes every output under ``oracle/_ref/`` (git-ignored, but shipped to the GPU box
by gpurun so the reference CPU and CUDA ops can be executed there as the parity arbiter and
the ``bench.py --impl reference`` arm).

Flags mirror the reference's build (setup.py:80-112): ``-std=c++17 -O3 -fopenmp
-DAT_PARALLEL_OPENMP=1 -DTORCH18 [-DWITH_CUDA]``; nvcc ``-O3 -DNDEBUG --expt-extended-lambda``.
The only addition is the target arch (the reference passes none, setup.py:101):
``-gencode arch=compute_100a,code=sm_100a``.

The one deviation from "unmodified"
-----------------------------------
torch >= 2.0 deleted ``TensorIteratorConfig::add_input(TensorBase&&)``; the reference passes
temporaries at 4 sites in ops/cpu/lsq_cpu.cpp and 8 in ops/cuda/lsq_cuda.cu.
Those two translation units are streamed through a purely mechanical rewrite (bind the two
``_unsafe_view`` temporaries to named lvalues) into ``oracle/_ref/_build/`` and compiled from
there; the arithmetic is untouched. The rewritten files are deleted after the build.

Outputs
-------
oracle/_ref/torchlsq/            reference Python package (copied verbatim) + _C.so
oracle/_ref/BUILD_INFO.json      what was built, with which flags, from which source hashes

Usage:  python oracle/build_ref.py [--cpu-only] [--force]
Nothing in the product (the ``torchlsq`` drop-in under lsqfakequantize-pytorch_b200/) imports
this directory.
"""
import argparse
import hashlib
import json
import os
import re
import shutil
import subprocess
import sys
import sysconfig
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
REF = Path(os.environ.get("LSQ_REFERENCE_ROOT", "/root/reference"))
OUT = HERE / "_ref"
BUILD = OUT / "_build"


def sh(cmd, **kw):
    print("+", " ".join(map(str, cmd)), flush=True)
    subprocess.check_call(list(map(str, cmd)), **kw)


def shim_d11(src: str) -> str:
    """Bind `torch::_unsafe_view(scale|shift, expected_shape)` temporaries to lvalues (D11)."""
    out = src.replace("torch::_unsafe_view(scale, expected_shape))", "scale_v)")
    out = out.replace("torch::_unsafe_view(shift, expected_shape))", "shift_v)")
    decl = ("\n    torch::Tensor scale_v = torch::_unsafe_view(scale, expected_shape);"
            "\n    torch::Tensor shift_v = torch::_unsafe_view(shift, expected_shape);")
    out, n = re.subn(r"(expected_shape\[axis\] = x\.size\(axis\);)", r"\1" + decl, out)
    assert n >= 2, "shim anchor not found"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-only", action="store_true")
    ap.add_argument("--force", action="store_true")
    args = ap.parse_args()

    if not REF.exists():
        print(f"[build_ref] {REF} not present (GPU box?) - using prebuilt oracle/_ref if any")
        return 0
    so = OUT / "torchlsq" / "_C.so"
    info_path = OUT / "BUILD_INFO.json"
    want_cuda = not args.cpu_only
    if so.exists() and info_path.exists() and not args.force:
        info = json.loads(info_path.read_text())
        if info.get("with_cuda") or not want_cuda:
            print("[build_ref] up to date:", so)
            return 0

    import torch
    from torch.utils import cpp_extension as ce

    csrc = REF / "torchlsq" / "csrc"
    if OUT.exists():
        shutil.rmtree(OUT)
    BUILD.mkdir(parents=True)
    # Python package, verbatim (import-time glue only; the product never imports it).
    shutil.copytree(REF / "torchlsq", OUT / "torchlsq",
                    ignore=shutil.ignore_patterns("csrc", "__pycache__", "*.so"))

    inc = [f"-I{p}" for p in ce.include_paths("cuda" if want_cuda else "cpu")]
    inc += [f"-I{sysconfig.get_paths()['include']}", f"-I{csrc}",
            # so the relative `#include "../global_scope.h"` of the two shimmed files resolves
            f"-I{csrc / 'ops' / 'cpu'}", f"-I{csrc / 'ops' / 'cuda'}"]
    defs = ["-DTORCH18", "-DTORCH_API_INCLUDE_EXTENSION_H", "-DTORCH_EXTENSION_NAME=_C",
            f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"]
    if want_cuda:
        defs.append("-DWITH_CUDA")
    cxx = ["g++", "-std=c++17", "-O3", "-fPIC", "-fopenmp", "-DAT_PARALLEL_OPENMP=1",
           "-Wno-unused-but-set-variable", "-Wno-unused-variable", "-Wno-sign-compare",
           "-Wno-unknown-pragmas", "-Wno-unused-function", "-Wno-deprecated-declarations"]

    cpu_shim = BUILD / "lsq_cpu_d11.cpp"
    cpu_shim.write_text(shim_d11((csrc / "ops/cpu/lsq_cpu.cpp").read_text()))
    cpp_sources = [csrc / "torchlsq.cpp", csrc / "ops/lsq.cpp",
                   csrc / "ops/autograd/lsq_autograd.cpp", cpu_shim]
    objs, procs = [], []
    t0 = time.time()
    for s in cpp_sources:
        o = BUILD / (s.stem + ".o")
        objs.append(o)
        cmd = cxx + defs + inc + ["-c", str(s), "-o", str(o)]
        print("+", " ".join(cmd), flush=True)
        procs.append(subprocess.Popen(cmd))
    if want_cuda:
        cu_shim = BUILD / "lsq_cuda_d11.cu"
        cu_shim.write_text(shim_d11((csrc / "ops/cuda/lsq_cuda.cu").read_text()))
        o = BUILD / "lsq_cuda_d11.o"
        objs.append(o)
        cmd = ["nvcc", "-std=c++17", "-O3", "-DNDEBUG", "--expt-extended-lambda",
               "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
               "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
               "-D__CUDA_NO_BFLOAT16_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__",
               "-Xcompiler", "-fPIC", "-Xcompiler", "-Wno-deprecated-declarations",
               "-diag-suppress", "20012,20013,20014,20015"] + defs + inc + \
              ["-c", str(cu_shim), "-o", str(o)]
        print("+", " ".join(cmd), flush=True)
        procs.append(subprocess.Popen(cmd))
    rc = [p.wait() for p in procs]
    if any(rc):
        print("[build_ref] compile failed", rc)
        return 1
    libs = [f"-L{p}" for p in ce.library_paths("cuda" if want_cuda else "cpu")]
    link = ["g++", "-shared", "-fopenmp", "-o", str(so)] + list(map(str, objs)) + libs + \
           ["-lc10", "-ltorch", "-ltorch_cpu", "-ltorch_python"]
    if want_cuda:
        link += ["-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    link += [f"-Wl,-rpath,{ce.library_paths('cpu')[0]}"]
    sh(link)
    sh(["strip", "--strip-unneeded", so])
    hashes = {str(p.relative_to(REF)): hashlib.sha256(p.read_bytes()).hexdigest()[:16]
              for p in sorted(csrc.rglob("*")) if p.is_file()}
    info_path.write_text(json.dumps({
        "with_cuda": want_cuda, "torch": torch.__version__, "seconds": round(time.time() - t0, 1),
        "cxx_flags": cxx[1:], "nvcc_arch": "compute_100a/sm_100a" if want_cuda else None,
        "shim": "D11 lvalue bind of _unsafe_view temporaries (lsq_cpu.cpp, lsq_cuda.cu)",
        "source_sha256_16": hashes}, indent=1))
    shutil.rmtree(BUILD)
    print(f"[build_ref] built {so} ({so.stat().st_size / 1e6:.1f} MB) in {time.time() - t0:.0f}s")
    return 0


if __name__ == "__main__":
    sys.exit(main())
