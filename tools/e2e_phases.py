"""Where the host-side time of one omniquant() block goes (dev tool)."""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
from types import SimpleNamespace
from omniquant_b200.main import quant_param_dicts
from omniquant_b200.quantize import omniquant as OQ
from omniquant_b200 import engine as E

marks = []
def mark(name):
    torch.cuda.synchronize(); marks.append((name, time.perf_counter()))

# monkey-patch phase functions to record time
for fn in ("fp_pass", "train", "requant_pass"):
    orig = getattr(E, fn)
    def wrap(*a, _o=orig, _n=fn, **k):
        mark(f"before {_n}"); r = _o(*a, **k); mark(f"after {_n}"); return r
    setattr(E, fn, wrap)
cls = OQ.QuantLlamaDecoderLayer
for meth in ("smooth_and_quant_inplace", "register_scales_and_zeros", "half", "float", "_ensure_engine"):
    orig = getattr(cls, meth)
    def wrap(self, *a, _o=orig, _n=meth, **k):
        mark(f"before {_n}"); r = _o(self, *a, **k); mark(f"after {_n}"); return r
    setattr(cls, meth, wrap)

W = bench.block_weights()
x_host = bench.calib_inputs(range(128)).pin_memory()
dev = torch.device("cuda")
a = SimpleNamespace(steps=2)
# reuse bench's e2e builder pieces
res = None
import bench as B
orig_omni = OQ.omniquant
def timed(*args, **kw):
    mark("omniquant start"); r = orig_omni(*args, **kw); mark("omniquant end"); return r
B_omni = timed
import omniquant_b200.quantize.omniquant as M
M.omniquant = timed
out = bench.run_e2e(SimpleNamespace(steps=2), W, x_host, dev, 1, 0)
print(out)
# print the phases of the last call
last = [i for i, (n, _) in enumerate(marks) if n == "omniquant start"][-1]
t0 = marks[last][1]
prev = t0
for n, t in marks[last:]:
    print(f"{n:40s} +{(t - prev) * 1e3:8.1f} ms   @{(t - t0) * 1e3:9.1f}")
    prev = t
