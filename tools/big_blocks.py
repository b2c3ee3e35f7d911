"""Full-size single blocks of BASELINE configs 3 and 4 on one B200 (dev tool): LLaMA-2-70B W2A16g64 LWC and
Falcon-180B W3A16g128 LWC.  A few calibration steps each: step time (CUDA graph), peak memory, finite decreasing loss."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from omniquant_b200 import engine as E  # noqa: E402


def weights(family, H, F_, heads, kv, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    r = lambda *s: (torch.randn(*s, generator=g, device="cuda") * 0.02).half().float()
    d = H // heads
    if family == "llama":
        return {"q": r(H, H), "k": r(kv * d, H), "v": r(kv * d, H), "o": r(H, H), "gate": r(F_, H), "up": r(F_, H),
                "down": r(H, F_), "ln1": torch.ones(H), "ln2": torch.ones(H)}
    return {"qkv": r((heads + 2 * kv) * d, H), "dense": r(H, H), "fc1": r(F_, H), "fc2": r(H, F_),
            "ln1": torch.ones(H), "ln2": torch.ones(H), "ln1.bias": torch.zeros(H), "ln2.bias": torch.zeros(H)}


def run(name, family, H, F_, heads, kv, wbits, group, nsamples=4, epochs=2, T=2048):
    torch.cuda.reset_peak_memory_stats()
    W = weights(family, H, F_, heads, kv)
    q = E.QuantConfig(wbits=wbits, abits=16, group_size=group, lwc=True, let=False)
    eng = E.BlockEngine(family, W, n_heads=heads, seq_len=T, qcfg=q, eps=1e-5, n_kv_heads=kv)
    del W
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(nsamples, T, H, generator=g, device="cuda").to(torch.bfloat16)
    fp, qi = x.clone(), x.clone()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = E.calibrate_block(eng, fp, qi, epochs=epochs)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    losses = res["losses"].float().cpu().reshape(epochs, -1).mean(1).tolist()
    # steady-state step
    tgt = fp[0]
    for _ in range(2):
        eng.train_step(x[0], tgt)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(4):
        eng.train_step(x[0], tgt)
    e1.record()
    torch.cuda.synchronize()
    n_par = sum(c for _, c in eng.layout.values())
    out = dict(block=name, step_ms=e0.elapsed_time(e1) / 4, wall_s=wall, epoch_mean_loss=losses,
               peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9, lwc_params=n_par,
               finite=bool(torch.isfinite(res["losses"]).all()), qout_finite=bool(torch.isfinite(qi.float()).all()))
    print(json.dumps(out), flush=True)
    del eng, x, fp, qi, res
    torch.cuda.empty_cache()


if __name__ == "__main__":
    which = sys.argv[1:] or ["70b", "falcon"]
    if "70b" in which:
        run("llama-2-70b W2A16g64 LWC", "llama", 8192, 28672, 64, 8, 2, 64)
    if "falcon" in which:
        run("falcon-180b W3A16g128 LWC", "falcon", 14848, 59392, 232, 8, 3, 128)
