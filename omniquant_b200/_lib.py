"""Loader / builder of the C-ABI shared library (``lib/liboq_b200.so``).

The ctypes signatures are generated from ``include/omniquant_b200.h`` so the
binding cannot drift from the declared ABI.  There is no fallback: if the
library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
import re
import shutil
import subprocess
from typing import Dict, List, Tuple

_PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_PKG)
HEADER = os.path.join(ROOT, "include", "omniquant_b200.h")
CSRC = os.path.join(_PKG, "csrc")
LIB_PATH = os.path.join(_PKG, "lib", "liboq_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O3", "-lineinfo",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
]

_CTYPES = {
    "int": ctypes.c_int, "long": ctypes.c_long, "float": ctypes.c_float, "oq_stream_t": ctypes.c_void_p,
}


def sources() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into one shared library (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and liboq_b200.so is missing/stale: cannot build the CUDA library")
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


def declared_functions() -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    """Parse the C header: name -> (return type, [(type, name)...])."""
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(int|const char\*)\s+(oq_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                mm = re.match(r"(.*?)(\w+)$", a)
                params.append((mm.group(1).strip(), mm.group(2)))
        out[name] = (ret, params)
    return out


def _ctype(t: str):
    if "*" in t:
        return ctypes.c_void_p
    return _CTYPES[t.replace("const ", "").strip()]


class OqError(RuntimeError):
    pass


class _Lib:
    def __init__(self, path: str):
        if not os.path.exists(path):
            raise OqError(f"{path} not found: build it with omniquant_b200._lib.build() (needs nvcc); "
                          "there is no CPU or PyTorch fallback for the CUDA kernels")
        self.path = path
        self.cdll = ctypes.CDLL(path)
        self.cdll.oq_last_error.restype = ctypes.c_char_p
        self.fns = {}
        self.launches = 0      # library calls issued (each enqueues one kernel; a few add a memset)
        for name, (ret, params) in declared_functions().items():
            fn = getattr(self.cdll, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = ctypes.c_char_p if "char" in ret else ctypes.c_int
            fn.argtypes = [_ctype(t) for t, _ in params]
            self.fns[name] = fn

    def call(self, name: str, *args):
        self.launches += 1
        rc = self.fns[name](*args)
        if rc != 0:
            msg = self.cdll.oq_last_error()
            raise OqError(f"{name} failed (rc={rc}): {msg.decode() if msg else ''}")


_LIB = None


def lib() -> _Lib:
    global _LIB
    if _LIB is None:
        # the built library ships in-tree; only (re)build here when it is absent or OQ_REBUILD=1
        if (not os.path.exists(LIB_PATH) or os.environ.get("OQ_REBUILD") == "1") and \
                (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
            build(force=True)
        _LIB = _Lib(LIB_PATH)
    return _LIB
