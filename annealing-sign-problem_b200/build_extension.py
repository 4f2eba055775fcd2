"""Builds ``libasp_b200.so`` (the C-ABI library) with nvcc for sm_100a and exposes the cffi
binding, in the style of the reference's annealing_sign_problem/build_extension.py:3-33
(``ffibuilder.cdef(<C declarations>)`` + a native build step).

The reference compiles one C file into a CPython extension (API mode).  Here the native
side is CUDA, so nvcc produces a plain shared library and cffi binds it in ABI mode
(``ffi.dlopen``) from the very declarations of ``include/asp_b200.h``.

    python annealing-sign-problem_b200/build_extension.py      # build in-tree
"""
import os
import re
import subprocess
import sys

from cffi import FFI

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HEADER = os.path.join(ROOT, "include", "asp_b200.h")
CSRC = os.path.join(HERE, "csrc")
LIBRARY = os.path.join(HERE, "libasp_b200.so")
SOURCES = ["operator.cu", "extract.cu", "extract_fused.cu", "legacy.cu", "reduce.cu", "anneal.cu", "greedy.cu", "apply.cu", "host.cu", "peer.cu", "sampling.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _declarations() -> str:
    """The C declarations of include/asp_b200.h, minus preprocessor lines, for cffi."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    lines = []
    for line in text.splitlines():
        stripped = line.strip()
        if stripped.startswith("#") or stripped.startswith('extern "C"') or stripped == "}":
            continue
        lines.append(line)
    body = "\n".join(lines)
    consts = "\n".join(
        "#define %s %s" % (m.group(1), m.group(2).strip("()"))
        for m in re.finditer(r"#define\s+(ASP_[A-Z_]+)\s+(\(?-?\d+\)?)", open(HEADER).read())
    )
    return consts + "\n" + body


ffibuilder = FFI()
ffibuilder.cdef(_declarations())


def needs_build() -> bool:
    if not os.path.exists(LIBRARY):
        return True
    built = os.path.getmtime(LIBRARY)
    deps = [HEADER] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> libasp_b200.so (in-tree)."""
    if not force and not needs_build():
        return LIBRARY
    cmd = ["nvcc"] + NVCC_FLAGS + ["-I", os.path.join(ROOT, "include"), "-o", LIBRARY]
    cmd += [os.path.join(CSRC, f) for f in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIBRARY


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print("built", LIBRARY)
