"""Builds libsyldet_cuda.so in-tree with nvcc for sm_100a (no torch, no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libsyldet_cuda.so")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "-ccbin", "g++"]

# (source, extra flags). kernels_generic.cu keeps IEEE operation order: no FMA contraction.
SOURCES = [
    ("config.cpp", []),
    ("capi.cpp", []),
    ("engine.cu", []),
    ("stream.cu", []),
    ("resample.cu", []),
    ("kernels_generic.cu", ["-fmad=false"]),
    ("kernels_fused.cu", ["-Xptxas", "-v"] + os.environ.get("SYLDET_FUSED_DEFS", "").split()),
    ("kernels_tc.cu", ["-Xptxas", "-v"] + os.environ.get("SYLDET_TC_DEFS", "").split()),
    ("kernels_wide.cu", ["-Xptxas", "-v"] + os.environ.get("SYLDET_WIDE_DEFS", "").split()),
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".hpp", ".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "syldet.h"))
    objs = []
    for src, extra in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [NVCC] + ARCH + COMMON + extra + ["-x", "cu", "-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for " + src)
    if force or _stale(OUT, objs):
        cmd = [NVCC] + ARCH + ["-shared", "-o", OUT] + objs + ["-cudart", "static", "-ccbin", "g++"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    cli_src = os.path.join(HERE, "..", "cli", "syldet_cli.cpp")
    cli_out = os.path.join(HERE, "syldet")
    if force or _stale(cli_out, [cli_src, OUT]):
        cmd = ["g++", "-O2", "-std=c++17", "-Wall", cli_src, "-o", cli_out, "-L" + HERE, "-lsyldet_cuda", "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    for name in ("syldet_stream_bench",):
        src = os.path.join(HERE, "..", "cli", name + ".cpp")
        out = os.path.join(HERE, name)
        if force or _stale(out, [src, OUT]):
            cmd = ["g++", "-O2", "-std=c++17", "-Wall", src, "-o", out, "-L" + HERE, "-lsyldet_cuda", "-Wl,-rpath,$ORIGIN"]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
