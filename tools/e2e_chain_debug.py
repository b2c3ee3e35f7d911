"""Diagnostic: omniquant() over a chained multi-block synthetic LLaMA-7B model (W4A4 LWC+LET) with per-epoch losses and
per-block stream statistics (OQ_DEBUG_STREAMS=1).  usage: python tools/e2e_chain_debug.py [blocks] [epochs] [nsamples]"""
import logging
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["OQ_DEBUG_STREAMS"] = "1"
import bench  # noqa: E402
from omniquant_b200.main import quant_param_dicts  # noqa: E402
from omniquant_b200.quantize.omniquant import omniquant  # noqa: E402

blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ns = int(sys.argv[3]) if len(sys.argv) > 3 else 128
c = bench.Cfg(3, epochs, ns)
dev = torch.device("cuda", 0)
W = block_w = bench.block_weights(c)
x = bench.calib_inputs(c, range(ns)).pin_memory()
args = SimpleNamespace(net=c.net, nsamples=ns, batch_size=1, epochs=c.epochs, let=True, lwc=True, aug_loss=False, alpha=0.5,
                       let_lr=5e-3, lwc_lr=1e-2, wd=0.0, wbits=4, abits=4, group_size=None, symmetric=False,
                       w_dynamic_method="per_channel", a_dynamic_method="per_token", resume=None, real_quant=False,
                       output_dir=None, cuda_graph=os.environ.get("OQ_NO_GRAPH") != "1", deactive_amp=True)
quant_param_dicts(args)
model, act, layers = bench._host_model(c, W, blocks)

# ---- per-step finiteness watch (OQ_NO_GRAPH=1): at the first non-finite loss dump the forward buffers of that step
from omniquant_b200 import engine as E  # noqa: E402
_orig_step = E.BlockEngine.train_step
_state = {"n": 0}


def _fin(t):
    t = t.codes if hasattr(t, "codes") else t
    return bool(torch.isfinite(t.float()).all())


def _watched(self, x, target, target2=None, allreduce=None):
    out = _orig_step(self, x, target, target2, allreduce)
    _state["n"] += 1
    if os.environ.get("OQ_NO_GRAPH") == "1" and not bool(torch.isfinite(out).all()):
        print(f"first non-finite loss at engine step {_state['n']} (AdamW step {int(self.step_dev)})")
        print("  input x finite", _fin(x), "target finite", _fin(target))
        st = lambda t: f"min {float(t.float().min()):.4g} max {float(t.float().max()):.4g} finite {_fin(t)}"
        for k in self.layout:
            print("  param", k, st(self.p[k]), "| grad", st(self.g[k]))
        for k, v in self.tmp.items():
            if isinstance(v, torch.Tensor):
                print("  tmp", k, st(v))
        for g, w in self.Wq.items():
            print("  Wq", g, st(w))
        for k, q in self.qstat.items():
            print("  qstat", k, "scale", st(q["scale"]), "zp", st(q["zp"]), "xmin", st(q["xmin"]), "xmax", st(q["xmax"]))
        order = ["a1", "a1q", "qkv", "qkvq", "S", "P", "O", "Oq", "h1", "a2", "a2q", "f1", "mq", "out", "dout"]
        for k in order + [k for k in self.bufs if k not in order]:
            b = self.bufs.get(k)
            if b is None:
                continue
            t = b.codes if hasattr(b, "codes") else b
            if not _fin(t):
                tf = t.float()
                bad = ~torch.isfinite(tf)
                rows = bad.reshape(-1, tf.shape[-1]).any(1).nonzero().flatten()[:5].tolist()
                print(f"  NON-FINITE {k} {tuple(t.shape)} nan {float(torch.isnan(tf).float().mean()):.4f} inf {float(torch.isinf(tf).float().mean()):.4f} first bad rows {rows}")
            else:
                print(f"  ok         {k} {tuple(t.shape)} absmax {float(t.float().abs().max()):.4g}")
        qv, qq = self.bufs.get("qkv"), self.bufs.get("qkvq")
        if qv is not None and qq is not None:
            d = self.d
            bad = (~torch.isfinite(qq.float())).view(qq.shape[0], -1, d).any(-1)       # [rows, heads3]
            idx = bad.nonzero()[:12].tolist()
            print("  NaN (row, head-slot) pairs:", idx, " per segment q/k/v:",
                  [int(bad[:, a:b].sum()) for a, b in ((0, self.h), (self.h, self.h + self.hkv), (self.h + self.hkv, self.h + 2 * self.hkv))])
            for r, h in idx[:6]:
                row = qv[r, h * d:(h + 1) * d].float()
                print(f"    row {r} slot {h}: raw qkv min {float(row.min()):.6g} max {float(row.max()):.6g} unique {int(row.unique().numel())} "
                      f"| a1q row absmax {float(self.bufs['a1q'][r].float().abs().max()):.4g} nonzero {int((self.bufs['a1q'][r] != 0).sum())} "
                      f"| a1 row absmax {float(self.bufs['a1'][r].float().abs().max()):.4g}")
        raise FloatingPointError("stopped at the first non-finite step")
    return out


E.BlockEngine.train_step = _watched
logging.basicConfig(level=logging.INFO, stream=sys.stdout, format="%(message)s")
lm = SimpleNamespace(model=model, device=dev, _device=dev, seqlen=bench.T)
try:
    omniquant(lm, args, None, act, None, logging.getLogger("oq"), inps=x)
except FloatingPointError as e:
    print("FAILED:", e)
    from omniquant_b200.models import _block
    torch.cuda.synchronize()
    for eng in ([] if os.environ.get("OQ_NO_GRAPH") == "1" else _block._ENGINES.values()):
        fin = lambda t: bool(torch.isfinite((t.codes if hasattr(t, "codes") else t).float()).all())
        st = lambda t: f"min {float(t.float().min()):.4g} max {float(t.float().max()):.4g} finite {fin(t)}"
        for k in eng.layout:
            print("  param", k, st(eng.p[k]), "| grad", st(eng.g[k]))
        for g, w in eng.Wq.items():
            print("  Wq", g, st(w), "| W", st(eng.W[g]))
        for k, st_ in eng.qstat.items():
            print("  qstat", k, "scale", st(st_["scale"]), "zp", st(st_["zp"]))
        for k, b in eng.bufs.items():
            if not fin(b):
                print("  non-finite buffer", k, tuple(b.shape), b.dtype, "nan frac", float(torch.isnan(b.float()).float().mean()),
                      "inf frac", float(torch.isinf(b.float()).float().mean()))
        print("  step", int(eng.step_dev))
