"""CPU-side checks of the product code: the C ABI library builds for sm_100a, loads and exports
every symbol include/omniquant_b200.h declares (no compute call without a GPU), argument errors are
reported through the ABI, and the host build of csrc/oq_math.h (the arithmetic the kernels inline)
reproduces the reference fixtures bit for bit."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from golden_util import ROOT, load

KAT = load("quantizer_kat.npz")
W_CASES = sorted({k.split("/")[0] for k in KAT.files if k.startswith("w") or k.startswith("kat")})


def test_library_exports_declared_abi():
    from omniquant_b200 import _lib
    _lib.build()
    decl = _lib.declared_functions()
    assert len(decl) >= 20
    l = _lib.lib()
    assert set(l.fns) == set(decl)
    assert l.cdll.oq_version() >= 100
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in decl:
        assert f" T {name}" in out, name


def test_bad_arguments_fail_loudly_without_gpu():
    from omniquant_b200 import _lib
    l = _lib.lib()
    with pytest.raises(_lib.OqError, match="n_bits"):
        l.call("oq_weight_quant_fwd", 16, 4, 4, None, None, 0, None, None, 1, 0, 0, None, None, 0, None, 0, None, None,
               None, None, None, None, None, None, None, 0, None, None, None, None)
    with pytest.raises(_lib.OqError, match="multiples of 8"):
        l.call("oq_gemm_bf16", 16, 0, 7, 0, 16, 0, 8, 0, 16, 0, 8, 0, 8, 8, 8, 1, None, None, 0, 0, 1.0, 0, 0, 0, 0, None)


def test_ops_refuse_cpu_tensors():
    import torch
    from omniquant_b200 import ops, _lib
    with pytest.raises(_lib.OqError, match="CUDA"):
        ops.weight_quant_fwd(torch.zeros(4, 4), n_bits=4)


@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("host") / "host_math.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host_math.cpp"), "-lm"])
    return ctypes.CDLL(so)


@pytest.mark.parametrize("case", W_CASES)
def test_product_arithmetic_matches_reference_fixtures(hostlib, case):
    g = lambda k: KAT[f"{case}/{k}"]
    nb, sym, grp, lwc = [int(v) for v in g("meta")]
    W = np.ascontiguousarray(g("W"), np.float32)
    N, K = W.shape
    G = -(-K // grp) if grp else 1
    if lwc and f"{case}/sig_up" in KAT.files:
        su, sl = np.ascontiguousarray(g("sig_up"), np.float32).ravel(), np.ascontiguousarray(g("sig_low"), np.float32).ravel()
    elif lwc:
        su = sl = np.full(N * G, np.float32(1.0) / (np.float32(1.0) + np.exp(np.float32(-4.0))), np.float32)
    else:
        su = sl = None
    deq, codes = np.empty((N, K), np.float32), np.empty((N, K), np.float32)
    s, z = np.empty(N * G, np.float32), np.empty(N * G, np.float32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p) if a is not None else None
    hostlib.host_fake_quant(P(W), N, K, grp, nb, sym, 0, P(su), P(sl), P(deq), P(s), P(z), P(codes))
    eq = lambda a, b: ((a.ravel().view(np.uint32) == np.ascontiguousarray(b, np.float32).ravel().view(np.uint32)) |
                       (np.isnan(a.ravel()) & np.isnan(np.asarray(b).ravel())) | ((a.ravel() == 0) & (np.asarray(b).ravel() == 0)))
    assert eq(s, g("scale")).all()
    assert eq(z, g("zp")).all()
    assert eq(deq, g("deq")).all()
    if f"{case}/codes" in KAT.files:
        cg = g("codes").reshape(N, -1)[:, :K]
        assert (np.isnan(cg) | (cg == codes)).all()


def test_truncate_and_fold(hostlib):
    hostlib.host_truncate.restype = ctypes.c_float
    hostlib.host_truncate.argtypes = [ctypes.c_float, ctypes.c_float]
    tr = lambda x: hostlib.host_truncate(x, 1e-2)
    assert tr(0.5) == 0.5 and tr(-0.004) == np.float32(-1e-2) and tr(0.0) == 0.0 and tr(0.003) == np.float32(1e-2)


def test_header_signatures_match_the_definitions():
    """ctypes binds from include/omniquant_b200.h; a parameter added in a .cu definition but not in the header (or the other
    way round) would be silent undefined behaviour.  Compare the parameter TYPE lists of every extern "C" definition under
    csrc/ with the declaration (cudaStream_t == oq_stream_t)."""
    import re
    from omniquant_b200 import _lib
    decl = _lib.declared_functions()
    norm = lambda t: " ".join(t.replace("cudaStream_t", "oq_stream_t").replace("const ", "const ").split())
    seen = {}
    for src in _lib.sources():
        text = open(src).read()
        text = re.sub(r"//[^\n]*", "", text)
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r'extern\s+"C"\s+(int|const char\*)\s+(oq_\w+)\s*\(([^)]*)\)\s*\{', text):
            ret, name, args = m.group(1), m.group(2), m.group(3).strip()
            types = []
            if args and args != "void":
                for a in args.split(","):
                    a = " ".join(a.split())
                    types.append(norm(re.match(r"(.*?)(\w+)$", a).group(1).strip()))
            seen[name] = (ret, types)
    assert set(seen) == set(decl), sorted(set(seen) ^ set(decl))
    for name, (ret, params) in decl.items():
        assert seen[name][0] == ret, name
        assert seen[name][1] == [norm(t) for t, _ in params], (name, seen[name][1], [t for t, _ in params])
