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
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC,-ffp-contract=off",
]
OBJ_DIR = os.path.join(_PKG, "lib", "obj")
HASH_PATH = LIB_PATH + ".srchash"

_CTYPES = {
    "int": ctypes.c_int, "long": ctypes.c_long, "float": ctypes.c_float, "oq_stream_t": ctypes.c_void_p,
}


def sources() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers() -> List[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh")))


def _digest(paths) -> str:
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()


def source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive the copy to the GPU box)."""
    return _digest(sources() + _headers() + [HEADER])


def _stale() -> bool:
    """True when the .so is missing or was built from other sources than the ones in the tree."""
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    return open(HASH_PATH).read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu under csrc/ for sm_100a into one shared library (in-tree): one object per source, compiled in
    parallel and re-used while its source and the shared headers are unchanged."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found and liboq_b200.so is missing/stale: cannot build the CUDA library")
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = _headers()
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        tag = _digest([src] + hdrs)
        tagf = obj + ".hash"
        if not force and os.path.exists(obj) and os.path.exists(tagf) and open(tagf).read() == tag:
            continue
        jobs.append((src, obj, tag, tagf))

    def compile_one(job):
        src, obj, tag, tagf = job
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode == 0:
            open(tagf, "w").write(tag)
        return src, res

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        for src, res in pool.map(compile_one, jobs):
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
            if verbose:
                print(res.stderr)
    objs = [os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o") for s in sources()]
    res = subprocess.run([nvcc, "-shared", "-o", LIB_PATH] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    open(HASH_PATH, "w").write(source_hash())
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
        # the built library ships in-tree; (re)build when it is absent, OQ_REBUILD=1, or built from other sources than the
        # ones in the tree (content hash) -- never run an old binary against a new header / Python side
        have_nvcc = bool(shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc"))
        if os.environ.get("OQ_REBUILD") == "1" and have_nvcc:
            build(force=True)
        elif _stale():
            if not have_nvcc:
                raise OqError(f"{LIB_PATH} is missing or stale (sources changed since it was built) and nvcc is not "
                              "available to rebuild it; there is no CPU or PyTorch fallback")
            build()
        _LIB = _Lib(LIB_PATH)
    return _LIB
