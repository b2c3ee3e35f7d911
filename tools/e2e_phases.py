"""Where the wall-clock time of one omniquant() block goes (dev tool): python tools/e2e_phases.py [n_blocks]."""
import sys
import time
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from omniquant_b200 import engine as E  # noqa: E402
from omniquant_b200 import streaming as S  # noqa: E402
from omniquant_b200.models import _block as BL  # noqa: E402

acc = {}


def wrap(obj, name, label=None):
    orig = getattr(obj, name)
    label = label or name

    def w(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = orig(*a, **k)
        torch.cuda.synchronize()
        acc[label] = acc.get(label, 0.0) + time.perf_counter() - t0
        return r
    setattr(obj, name, w)


for fn in ("fp_pass", "train", "requant_pass"):
    wrap(E, fn)
for m in ("fetch", "prefetch", "writeback", "drain"):
    wrap(S.BlockStreamer, m, "streamer." + m)
for m in ("smooth_and_quant_inplace", "register_scales_and_zeros", "_ensure_engine"):
    if hasattr(BL.QuantDecoderLayerBase, m):
        wrap(BL.QuantDecoderLayerBase, m, "layer." + m)

c = bench.Cfg(3)
n_blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 4
W = bench.block_weights(c)
x_host = bench.calib_inputs(c, range(c.nsamples)).pin_memory()
a = SimpleNamespace(steps=2 * n_blocks, e2e_blocks=n_blocks)
dev = torch.device("cuda")
# first call = warm-up inside run_e2e; reset the accumulators after it by wrapping omniquant
import omniquant_b200.quantize.omniquant as OQ  # noqa: E402
orig = OQ.omniquant
calls = []


def timed(*args, **kw):
    acc.clear()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = orig(*args, **kw)
    torch.cuda.synchronize()
    calls.append((time.perf_counter() - t0, dict(acc)))
    return r


OQ.omniquant = timed
out = bench.run_e2e(a, c, W, x_host, dev, 1, 0)
print(out["value"], "s/block over", out["steps"], "blocks")
tot, ph = calls[-1]
print(f"last call: {tot:.3f} s for {n_blocks} blocks = {tot / n_blocks:.3f} s/block")
for k, v in sorted(ph.items(), key=lambda kv: -kv[1]):
    print(f"  {k:36s} {v / n_blocks * 1e3:9.1f} ms/block")
print(f"  {'unaccounted':36s} {(tot - sum(ph.values())) / n_blocks * 1e3:9.1f} ms/block")
