"""One LLaMA-7B calibration step bracketed by cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import sys

import torch

sys.path.insert(0, ".")
from omniquant_b200.engine import BlockEngine, QuantConfig  # noqa: E402

wbits, abits, group, let = ([int(v) for v in sys.argv[1:5]] + [4, 4, 0, 1])[:4] if len(sys.argv) > 1 else (4, 4, 0, 1)
H, F_, heads, T = 4096, 11008, 32, 2048
g = torch.Generator().manual_seed(0)
r = lambda *s: (torch.randn(*s, generator=g) * 0.02).half()
W = {"q": r(H, H), "k": r(H, H), "v": r(H, H), "o": r(H, H), "gate": r(F_, H), "up": r(F_, H), "down": r(H, F_),
     "ln1": torch.ones(H), "ln2": torch.ones(H)}
eng = BlockEngine("llama", W, n_heads=heads, seq_len=T, qcfg=QuantConfig(wbits=wbits, abits=abits, group_size=group, lwc=True, let=bool(let)), eps=1e-6)
x = torch.randn(T, H, device="cuda").to(torch.bfloat16)
tgt = torch.randn(T, H, device="cuda").to(torch.bfloat16)
for _ in range(2):
    eng.train_step(x, tgt)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.train_step(x, tgt)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
