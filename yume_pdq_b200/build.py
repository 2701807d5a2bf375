"""Build libyume_pdq_cuda.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libyume_pdq_cuda.so")
# the reference's library name (Cargo.toml: crate `yume_pdq`, cdylib -> libyume_pdq.so): existing users link `-lyume_pdq`
# or load it through $LIBYUME_PDQ_PATH (integration/hash.py:32) and find this implementation
LIB_ALIAS = os.path.join(HERE, "libyume_pdq.so")
CLI_SRC = os.path.join(HERE, "cli")
CLI = os.path.join(HERE, "yume-pdq-cuda")     # `pipe` front end (cli/yume_pdq_pipe.cpp), links only the C ABI
SOURCES = ["api.cu", "hasher.cu", "matcher.cu", "matcher_tc.cu", "matcher_tc4.cu", "matcher_tc4x2.cu", "matcher_tc1.cu", "synth.cu", "peak.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps += [os.path.join(HERE, "..", "include", f) for f in ("yume_pdq.h", "yume_pdq_cuda.h", "yume_pdq_luts.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(tag: str, defines: list[str]) -> str:
    """Developer builds for kernel experiments (tools/quick_bench.py, YUME_PDQ_CUDA_LIB): the same sources with
    extra -D flags into libyume_pdq_cuda_<tag>.so.  Not used by the product or the tests."""
    out = os.path.join(HERE, f"libyume_pdq_cuda_{tag}.so")
    cmd = [_nvcc()] + NVCC_FLAGS + [f"-D{d}" for d in defines] + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    return out


def build_cli(force: bool = False) -> str:
    srcs = [os.path.join(CLI_SRC, f) for f in os.listdir(CLI_SRC)]
    deps = srcs + [LIB, os.path.join(HERE, "..", "include", "yume_pdq_cuda.h")]
    if not force and os.path.exists(CLI) and all(os.path.getmtime(d) <= os.path.getmtime(CLI) for d in deps):
        return CLI
    cmd = [shutil.which("g++") or "g++", "-O2", "-std=c++17", "-Wall", "-pthread",
           os.path.join(CLI_SRC, "yume_pdq_pipe.cpp"), "-L" + HERE, "-lyume_pdq_cuda", "-Wl,-rpath,$ORIGIN", "-o", CLI]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed (yume-pdq-cuda)")
    return CLI


def link_alias() -> None:
    if os.path.islink(LIB_ALIAS) and os.readlink(LIB_ALIAS) == os.path.basename(LIB):
        return
    if os.path.lexists(LIB_ALIAS):
        os.remove(LIB_ALIAS)
    os.symlink(os.path.basename(LIB), LIB_ALIAS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        link_alias()
        build_cli()
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + log)
    if r.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed")
    if verbose:
        print(log)
    link_alias()
    build_cli(force=True)
    return LIB


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "variant":      # python -m yume_pdq_b200.build variant <tag> <DEFINE>...
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        build(force=True, verbose=True)
