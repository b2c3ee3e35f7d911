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
logging.basicConfig(level=logging.INFO, stream=sys.stdout, format="%(message)s")
lm = SimpleNamespace(model=model, device=dev, _device=dev, seqlen=bench.T)
try:
    omniquant(lm, args, None, act, None, logging.getLogger("oq"), inps=x)
except FloatingPointError as e:
    print("FAILED:", e)
    from omniquant_b200.models import _block
    torch.cuda.synchronize()
    for eng in _block._ENGINES.values():
        fin = lambda t: bool(torch.isfinite(t.float()).all())
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
