#!/usr/bin/env python
"""Benchmark of the calibration hot path: LLaMA-7B W4A4 LWC+LET, seconds per decoder block.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is the calibration of ONE decoder block exactly as quantize/omniquant.py:157-250 does it:
FP pass over the 128 calibration samples, 20 epochs of AdamW steps on the LWC+LET parameters (global
batch = N, one sample per GPU per step, 128/N steps per epoch), in-place fold + quantise, and the
re-quantised forward pass over the 128 samples.  Random-init LLaMA-7B block weights (fp16 values),
synthetic calibration activations and LET statistics (no network in this environment).

`value` is device-timed (CUDA events, max over ranks) with weights and calibration activations already
resident in HBM; `e2e` is the same block through the public API (`omniquant()` on a one-block model)
with the block weights coming from pinned host memory, the quantised block going back to the host and
the calibration activations copied from pinned host memory, all inside the timed region.

--impl reference: the reference algorithm on the host cores (the oracle port, oracle/oq_oracle.py; the
reference is Python/PyTorch and cannot travel to the GPU box), a bounded sample of the same workload
(K optimiser steps + FP forwards at the same shapes) extrapolated to one block and labelled as such.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, FFN, HEADS, T = 4096, 11008, 32, 2048
NSAMPLES, EPOCHS = 128, 20
WBITS, ABITS = 4, 4
METRIC = "llama7b_w4a4_lwc_let_calib_s_per_block"
WORKLOAD = "LLaMA-7B decoder block, W4A4 per-channel weights / per-token activations, LWC+LET, 128x2048 synthetic tokens, 20 epochs"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1388.8), d.get("hbm_gbs", 6539.5), "measured"
    return 1400.0, 6650.0, "fallback"


def gemm_traffic_from_ncu():
    """Per-launch DRAM bytes (read + write) of the dominant kernel from the committed `ncu --set full` capture
    (profiles/r1_gemm_ncu_full_summary.csv, QuantLinear launches of gemm_kernel<256>), or None."""
    import csv
    path = os.path.join(ROOT, "profiles", "r1_gemm_ncu_full_summary.csv")
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    name = [i for i, h in enumerate(hdr) if h.startswith("Kernel Name")][0]
    rd = [i for i, h in enumerate(hdr) if h.startswith("dram__bytes_read.sum")][0]
    wr = [i for i, h in enumerate(hdr) if h.startswith("dram__bytes_write.sum")][0]
    vals = [(float(r[rd]) + float(r[wr])) * 1e6 for r in rows[1:] if "gemm_kernel<256>" in r[name] and float(r[wr]) < 200]
    return round(sum(vals) / len(vals)) if vals else None


def block_weights(seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: (torch.randn(*s, generator=g) * 0.02).half()
    W = {"q": r(H, H), "k": r(H, H), "v": r(H, H), "o": r(H, H), "gate": r(FFN, H), "up": r(FFN, H), "down": r(H, FFN),
         "ln1": torch.ones(H).half(), "ln2": torch.ones(H).half()}
    return W


def calib_inputs(indices, seed=1234):
    """Synthetic block inputs for the given global sample indices (one generator per sample, so every rank
    builds only its own shard and the data does not depend on the world size)."""
    out = torch.empty(len(indices), T, H, dtype=torch.bfloat16)
    for i, j in enumerate(indices):
        g = torch.Generator().manual_seed(seed + int(j))
        x = torch.randn(T, H, generator=g, dtype=torch.float32)
        x[:, ::512] *= 8.0          # a few outlier channels, what LET exists for
        out[i] = x.to(torch.bfloat16)
    return out


def let_init(W, seed=1234):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for key, wn in (("qkv", "q"), ("out", "o"), ("fc1", "up")):
        act = (torch.randn(H, generator=g).abs() * 2 + 1e-5).clamp(min=1e-5)
        w = W[wn].float().max(dim=0)[0].clamp(min=1e-5)
        out[f"{key}_s"] = (act.pow(0.5) / w.pow(0.5)).clamp(min=1e-5)
    return out


class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- ours
def run_ours(a):
    import torch.distributed as dist
    from omniquant_b200 import ops
    from omniquant_b200.engine import BlockEngine, QuantConfig, calibrate_block

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line (rank 0): everything libraries print (NCCL's version banner ...) goes to stderr
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == a.gpus, f"--gpus {a.gpus} but WORLD_SIZE={world} (launch with torch.distributed.run)"
    assert NSAMPLES % world == 0
    allreduce = (lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM)) if world > 1 else None

    W_host = block_weights()
    W = {k: v.to(dev) for k, v in W_host.items()}       # resident in HBM before the timed region
    n_local = NSAMPLES // world
    # rank r owns samples r, r+world, ... (global step j = samples [j*world, (j+1)*world), reference --batch_size world)
    x_host = calib_inputs(range(rank, NSAMPLES, world)).pin_memory()
    x_dev = x_host.to(dev)
    q = QuantConfig(wbits=WBITS, abits=ABITS, group_size=0, lwc=True, let=True, recip_div=True)
    init = let_init(W_host)

    resident = {}

    def fresh_engine():
        return BlockEngine("llama", W, n_heads=HEADS, seq_len=T, qcfg=q, eps=1e-6, device=dev, batch=1, let_init=init,
                           world_size=world)

    def new_engine():
        """The engine of the next block: like omniquant() on a multi-block model, the resident engine is re-loaded with
        the block's weights (copied + converted inside the timed region) and keeps the captured CUDA graph of the step."""
        if "eng" not in resident:
            resident["eng"] = fresh_engine()
            return resident["eng"]
        return resident["eng"].reload(W, init)

    def one_block(eng):
        fp, qi = x_dev.clone(), x_dev.clone()
        return calibrate_block(eng, fp, qi, epochs=EPOCHS, allreduce=allreduce, use_graph=True)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        one_block(new_engine())
    sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for _ in range(a.steps):
        last = one_block(new_engine())
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_block = float(ms) / a.steps
    launches_eager = ops.launch_count() - l0
    # the captured step replays its kernels from a CUDA graph: count them per replay
    steps_per_block = EPOCHS * n_local
    loss_mean = float(last["losses"].mean())
    loss_first_epoch, loss_last_epoch = float(last["losses"][:n_local].mean()), float(last["losses"][-n_local:].mean())

    # ---- roofline of the dominant kernel (gemm_kernel<256>: the QuantLinear fwd/dgrad/wgrad GEMMs), timed live
    eng = fresh_engine()
    eng.overlap = False      # time the GEMM launches alone on the stream (no quantiser kernels co-running beside them)
    xs, ts = x_dev[0], x_dev[1 % n_local]
    for _ in range(2):
        eng.train_step(xs, ts)
    torch.cuda.synchronize()
    ops.GEMM_TIMER = []
    l1 = ops.launch_count()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(3):
        eng.train_step(xs, ts)
    s1.record()
    launches_per_step = (ops.launch_count() - l1) // 3
    torch.cuda.synchronize()
    step_ms_eager = s0.elapsed_time(s1) / 3
    lin = [(e0_.elapsed_time(e1_), fl) for e0_, e1_, fl, batched in ops.GEMM_TIMER if not batched]
    att = [(e0_.elapsed_time(e1_), fl) for e0_, e1_, fl, batched in ops.GEMM_TIMER if batched]
    ops.GEMM_TIMER = None
    peak_tf, peak_hbm, which = peaks()
    lin_ms, lin_fl = sum(t for t, _ in lin), sum(f for _, f in lin)
    achieved = lin_fl / (lin_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "gemm_kernel<256> (QuantLinear fwd/dgrad/wgrad)", "achieved": round(achieved, 1),
                "peak": peak_tf, "unit": "TFLOP/s", "frac": round(achieved / peak_tf, 4), "traffic": gemm_traffic_from_ncu(),
                "traffic_note": "mean DRAM read+write bytes per QuantLinear launch, ncu --set full (profiles/r1_gemm_ncu_full_summary.csv); "
                                "algorithmic operand+result bytes per launch average 1.7e8",
                "peak_source": f"bf16_tflops_sustained, {which}", "launches_timed": len(lin),
                "timing": "CUDA events around each launch on the launching stream, 3 eager optimiser steps after the timed region, side-stream overlap off",
                "avg_launch_ms": round(lin_ms / max(len(lin), 1), 4),
                "attention_gemm_tflops": round(sum(f for _, f in att) / (sum(t for t, _ in att) * 1e-3) / 1e12, 1) if att else None}
    # whole-step tensor utilisation: algorithmic flops of one step (causal attention) / step time
    f_step = 3 * 2 * T * (4 * H * H + 3 * H * FFN) + 3 * 4 * HEADS * T * T * (H // HEADS) * 0.5
    f_block = steps_per_block * world * f_step + 2 * NSAMPLES * (f_step / 3)
    roofline["block_tflops"] = round(f_block / (ms_per_block * 1e-3) / 1e12 / world, 1)
    roofline["block_frac_of_peak"] = round(roofline["block_tflops"] / peak_tf, 4)
    del eng

    # ---- e2e through the public API with host buffers
    e2e = None
    if a.e2e:
        e2e = run_e2e(a, W_host, x_host, dev, world, rank)

    cpu = None
    if rank == 0 and world == 1 and a.cpu_baseline:
        cpu = cpu_baseline(steps=1)

    if rank == 0:
        out = {"metric": METRIC, "value": round(ms_per_block / 1e3, 4), "unit": "s/block", "n_gpus": world, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": round(ms_per_block, 2), "higher_is_better": False, "scaling": "strong",
               "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": WORKLOAD, "step": "one decoder block calibrated (FP pass + 20x128 AdamW steps + requant pass)",
                          "global_batch": world, "seq_len": T, "nsamples": NSAMPLES, "epochs": EPOCHS,
                          "parallelism": f"sample-sharded dp{world}",
                          "engine": "resident across blocks: each block's weights are copied + converted into the same device "
                                    "buffers inside the timed region, the optimiser-step CUDA graph is captured once (warm-up) and replayed",
                          "l2": "working set per optimiser step (weights 1.2 GB fp32 + activations) exceeds the 126 MB L2"},
               "optimizer_step_ms_eager_instrumented": round(step_ms_eager, 3), "launches_per_optimizer_step": launches_per_step,
               "loss_mean": loss_mean, "loss_first_epoch": loss_first_epoch, "loss_last_epoch": loss_last_epoch, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
               "gpu_launches": int(launches_eager + a.steps * max(steps_per_block - 3, 0) * launches_per_step),
               "clocks": clocks}
        json_out.write(json.dumps(out) + "\n")
        json_out.flush()
    # teardown: the captured step graphs hold the NCCL all-reduce kernel; release them BEFORE the communicator goes away
    # (destroying the process group under a live graph blocks), and never let teardown outlive the measurement
    import gc
    for e in resident.values():
        e._graph_state = None
    resident.clear()
    from omniquant_b200.models._block import release_engines
    release_engines()
    eng = last = None
    gc.collect()
    torch.cuda.synchronize()
    if world > 1:
        watchdog = threading.Timer(30.0, lambda: os._exit(0))
        watchdog.daemon = True
        watchdog.start()
        dist.barrier()
        dist.destroy_process_group()
        watchdog.cancel()


class _StridedHostInputs:
    """Duck-typed stand-in for the global calibration tensor on a rank that only holds its own shard in pinned
    host memory: omniquant() calls .to(device, dtype) and then indexes the rank's rows."""

    def __init__(self, shard, rank, world):
        self.shard, self.rank, self.world = shard, rank, world

    def to(self, dev, dtype):
        return _StridedDeviceInputs(self.shard.to(dev, dtype, non_blocking=True), self.rank, self.world)


class _StridedDeviceInputs:
    def __init__(self, shard, rank, world):
        self.shard, self.rank, self.world = shard, rank, world

    def __getitem__(self, idx):
        local = [(i - self.rank) // self.world for i in idx]
        assert all(i % self.world == self.rank for i in idx)
        return self.shard[local]


def run_e2e(a, W, x_host, dev, world, rank):
    """One block through omniquant() (the call a user makes): weights from pinned host memory, calibration
    activations from pinned host memory, quantised block + learned parameters back on the host."""
    from types import SimpleNamespace
    from transformers import LlamaConfig
    from omniquant_b200.main import quant_param_dicts
    from omniquant_b200.quantize.omniquant import omniquant

    cfg = LlamaConfig(hidden_size=H, intermediate_size=FFN, num_attention_heads=HEADS, num_key_value_heads=HEADS,
                      num_hidden_layers=1, vocab_size=32, max_position_embeddings=T, rms_norm_eps=1e-6)
    cfg.use_cache = False

    class Lin(torch.nn.Module):
        def __init__(self, w):
            super().__init__()
            self.weight = torch.nn.Parameter(w.clone().pin_memory(), requires_grad=False)
            self.bias = None
            self.out_features, self.in_features = w.shape

    class Norm(torch.nn.Module):
        def __init__(self, w):
            super().__init__()
            self.weight = torch.nn.Parameter(w.clone().pin_memory(), requires_grad=False)
            self.variance_epsilon = 1e-6

    def make_layer():
        layer = torch.nn.Module()
        layer.self_attn = torch.nn.Module()
        for n in ("q", "k", "v", "o"):
            setattr(layer.self_attn, f"{n}_proj", Lin(W[n]))
        layer.mlp = torch.nn.Module()
        for n in ("gate", "up", "down"):
            setattr(layer.mlp, f"{n}_proj", Lin(W[n]))
        layer.input_layernorm, layer.post_attention_layernorm = Norm(W["ln1"]), Norm(W["ln2"])
        return layer

    g = torch.Generator().manual_seed(1234)
    act = {f"model.layers.0.{n}": torch.randn(H, generator=g).abs() * 2 + 1e-5
           for n in ("self_attn.q_proj", "self_attn.o_proj", "mlp.up_proj")}
    args = SimpleNamespace(net="llama-7b", nsamples=x_host.shape[0] * world, batch_size=world, epochs=EPOCHS, let=True,
                           lwc=True, aug_loss=False, alpha=0.5, let_lr=5e-3, lwc_lr=1e-2, wd=0.0, wbits=WBITS, abits=ABITS,
                           group_size=None, symmetric=False, w_dynamic_method="per_channel", a_dynamic_method="per_token",
                           resume=None, real_quant=False, output_dir=None, cuda_graph=True)
    quant_param_dicts(args)

    # omniquant() shards `inps` by rank itself and only touches this rank's rows: hand it a pinned host view of the
    # global [nsamples, T, H] tensor in which the rank's own samples sit at stride `world` (other rows unused)
    if world == 1:
        inps_view = x_host
    else:
        inps_view = _StridedHostInputs(x_host, rank, world)

    def one():
        model = torch.nn.Module()
        model.config = cfg
        model.model = torch.nn.Module()
        model.model.layers = torch.nn.ModuleList([make_layer()])
        model.model.embed_tokens = torch.nn.Identity()
        model.model.norm = torch.nn.Identity()
        lm = SimpleNamespace(model=model, device=dev, _device=dev, seqlen=T)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        omniquant(lm, args, None, act, None, None, inps=inps_view)
        w0 = model.model.layers[0].self_attn.q_proj.weight
        assert w0.device.type == "cpu"
        torch.cuda.synchronize()
        return time.perf_counter() - t0

    one()                                       # warm-up
    n = max(1, a.steps // 2)
    import torch.distributed as dist
    ts = []
    for _ in range(n):
        if world > 1:
            dist.barrier()
        ts.append(one())
    t = torch.tensor([sum(ts) / n], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wbytes = sum(v.numel() * 2 for v in W.values())
    return {"value": round(float(t), 4), "unit": "s/block", "h2d_bytes_per_step": int(wbytes + x_host.numel() * 2),
            "d2h_bytes_per_step": int(wbytes + 4 * (2 * 7 * H + 7 * H)), "steps": n,
            "api": "omniquant_b200.quantize.omniquant.omniquant() on a one-block model; wall clock, max over ranks"}


# ---------------------------------------------------------------------------------------------- CPU reference arm
def cpu_baseline(steps=1):
    """The oracle port (torch fp32 on the host cores) at the same shapes: `steps` optimiser steps + one FP
    forward, extrapolated to a block: 2560*t_step + 256*t_fwd."""
    from oracle import oq_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    W = {k: v.float() for k, v in block_weights().items()}
    cfg = orc.QCfg(WBITS, ABITS, None, False, True, True)
    P = orc.init_lwc({k: W[k] for k in orc.LLAMA_LINEARS}, cfg)
    for k, v in let_init(W).items():
        P[k] = v.clone()
    for k in ("qkv_d", "out_d", "fc1_d"):
        P[k] = torch.zeros(H)
    P["qkt_s"] = torch.ones(H)
    x = calib_inputs(range(2)).float()
    cos, sin = orc.rope_tables(H // HEADS, T)
    mask = orc.causal_mask(T)
    with torch.no_grad():
        Tfp = orc.llama_prepare(W, P, orc.QCfg(WBITS, ABITS, None, False, True, False), quant=False)
        t0 = time.perf_counter()
        fp = orc.llama_forward(x[:1], Tfp, HEADS, cos, sin, mask, cfg, False)
        t_fwd = time.perf_counter() - t0
    keys = [k for k in P if k.endswith(("_s", "_d", ".up", ".low"))]
    for k in keys:
        P[k].requires_grad_(True)
    opt = torch.optim.AdamW([{"params": [P[k] for k in keys if "_" in k and "." not in k], "lr": 5e-3},
                             {"params": [P[k] for k in keys if "." in k], "lr": 1e-2}], weight_decay=0)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        Tq = orc.llama_prepare(W, P, cfg)
        out = orc.llama_forward(x[:1], Tq, HEADS, cos, sin, mask, cfg, True)
        loss = torch.nn.functional.mse_loss(fp, out)
        opt.zero_grad()
        loss.backward()
        opt.step()
        ts.append(time.perf_counter() - t0)
    t_step = sum(ts) / len(ts)
    block = EPOCHS * NSAMPLES * t_step + 2 * NSAMPLES * t_fwd
    return {"value": round(block, 1), "unit": "s/block", "cores": cores, "kind": "port",
            "sample": f"{steps} optimiser step(s) ({t_step:.2f} s each) + 1 FP forward ({t_fwd:.2f} s) at LLaMA-7B block shapes, "
                      f"fp32 torch on the host, extrapolated: {EPOCHS * NSAMPLES}*t_step + {2 * NSAMPLES}*t_fwd"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    cpu = cpu_baseline(steps=max(1, a.steps))
    out = {"impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "s/block", "n_gpus": a.gpus, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": round(cpu["value"] * 1e3, 1), "higher_is_better": False, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "note": "reference algorithm on the host cores (oracle port), bounded sample extrapolated to one block"},
           "cpu_baseline": cpu, "e2e": {"value": cpu["value"], "unit": "s/block", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 1)}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-e2e", dest="e2e", action="store_false")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
