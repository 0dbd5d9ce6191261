"""In-tree native build driver.

Two shared objects are produced under ``dgl_b200/lib``:

* ``libdglb200_cpu.so``  - host runtime (OpenMP kernels, samplers, hash maps,
  cache policies, socket RPC transport, shared memory, partitioners).
* ``libdglb200_cuda.so`` - every sm_100a CUDA kernel (compiled with
  ``-gencode arch=compute_100a,code=sm_100a -lineinfo``).

Both export a plain C ABI consumed through ``ctypes`` (see ``loader.py``), so
no libtorch headers are needed and a full rebuild takes seconds, not minutes.
This replaces the reference's four libraries + TVM packed-func FFI
(/root/reference/CMakeLists.txt:245-550, python/dgl/_ffi) with a single thin
ABI: tensors are allocated by PyTorch and passed as raw pointers + the current
CUDA stream handle.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "csrc"
LIBDIR = ROOT / "lib"
OBJDIR = ROOT / "lib" / "obj"

CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
CXX = os.environ.get("DGLB200_CXX", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")

CPU_FLAGS = ["-O3", "-std=c++17", "-fPIC", "-fopenmp", "-march=x86-64-v3", "-Wall",
             "-Wno-unused-function", "-Wno-sign-compare", "-Wno-unknown-pragmas", "-pthread"]
CUDA_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "--expt-relaxed-constexpr", "--expt-extended-lambda", "-Xcompiler", "-fPIC",
              "-ccbin", CXX, "-Xptxas", "-v", "-diag-suppress", "177,550"]


def _hash_inputs(src: Path, flags, headers) -> str:
    h = hashlib.sha1()
    h.update(src.read_bytes())
    h.update(" ".join(flags).encode())
    for hd in headers:
        h.update(hd.read_bytes())
    return h.hexdigest()


def _compile_one(job):
    src, obj, cmd, stamp, digest, logf = job
    if obj.exists() and stamp.exists() and stamp.read_text() == digest:
        return src, True, "cached"
    p = subprocess.run(cmd, capture_output=True, text=True)
    logf.write_text(p.stdout + p.stderr)
    if p.returncode != 0:
        return src, False, p.stdout + p.stderr
    stamp.write_text(digest)
    return src, True, p.stderr


def _build_lib(name, srcs, compiler_cmd, flags, headers, link_cmd, verbose):
    OBJDIR.mkdir(parents=True, exist_ok=True)
    jobs = []
    objs = []
    for src in srcs:
        obj = OBJDIR / (name + "_" + src.stem + ".o")
        stamp = OBJDIR / (name + "_" + src.stem + ".stamp")
        logf = OBJDIR / (name + "_" + src.stem + ".log")
        digest = _hash_inputs(src, flags, headers)
        cmd = compiler_cmd + flags + ["-I", str(CSRC), "-c", str(src), "-o", str(obj)]
        jobs.append((src, obj, cmd, stamp, digest, logf))
        objs.append(obj)
    nthreads = int(os.environ.get("DGLB200_BUILD_JOBS", max(2, (os.cpu_count() or 4) // 2)))
    changed = False
    with ThreadPoolExecutor(nthreads) as ex:
        for src, ok, msg in ex.map(_compile_one, jobs):
            if not ok:
                raise RuntimeError(f"compile failed: {src}\n{msg}")
            if msg != "cached":
                changed = True
                if verbose:
                    print(f"[build] compiled {src.name}", file=sys.stderr)
    out = LIBDIR / f"lib{name}.so"
    if changed or not out.exists():
        tmp = LIBDIR / f"lib{name}.so.tmp{os.getpid()}"
        cmd = link_cmd + [str(o) for o in objs] + ["-o", str(tmp)]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed: {name}\n{p.stdout}{p.stderr}")
        os.replace(tmp, out)
        if verbose:
            print(f"[build] linked {out}", file=sys.stderr)
    return out


def build_cpu(verbose=True):
    srcs = sorted((CSRC / "cpu").glob("*.cc"))
    headers = sorted((CSRC / "cpu").glob("*.h")) + sorted(CSRC.glob("*.h"))
    return _build_lib("dglb200_cpu", srcs, [CXX], CPU_FLAGS, headers,
                      [CXX, "-shared", "-fopenmp", "-pthread", "-lrt"], verbose)


def have_nvcc():
    return os.path.exists(NVCC)


def build_cuda(verbose=True):
    if not have_nvcc():
        raise RuntimeError("nvcc not found at " + NVCC)
    srcs = sorted((CSRC / "cuda").glob("*.cu"))
    headers = sorted((CSRC / "cuda").glob("*.cuh")) + sorted((CSRC / "cuda").glob("*.h")) + sorted(CSRC.glob("*.h"))
    return _build_lib("dglb200_cuda", srcs, [NVCC], CUDA_FLAGS, headers,
                      [NVCC, "-shared", "-cudart", "static", "-Xcompiler", "-fPIC",
                       "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", CXX], verbose)


def build_all(verbose=True):
    LIBDIR.mkdir(parents=True, exist_ok=True)
    outs = [build_cpu(verbose)]
    if have_nvcc():
        outs.append(build_cuda(verbose))
    return outs


def clean():
    if OBJDIR.exists():
        shutil.rmtree(OBJDIR)
    for f in LIBDIR.glob("*.so"):
        f.unlink()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "clean":
        clean()
    else:
        for o in build_all():
            print(o)
