"""Block calibration engine: one decoder block's LWC/LET training step as an explicit,
hand-scheduled sequence of sm_100a kernels (forward, backward, optimiser), no autograd.

This is the B200-native replacement for the body of the reference's per-block loop
(quantize/omniquant.py:157-250): ``smooth_and_quant_temporary`` -> block forward -> MSE ->
backward -> AdamW, plus the FP pass (:165-172), the in-place fold/quantise (:239) and the
re-quantised forward pass (:242-245).  The reference's module API (omniquant_b200/quantize,
omniquant_b200/models) drives this engine; all device compute is the C-ABI library.

Data layout in HBM (M = batch*T token rows, bf16 activations, fp32 masters):
  weights   fp32 masters, concatenated by consumer:  Wqkv [3H,H], Wo [H,H], Wgu [2F,H] (gate|up) or
            Wfc1 [F,H], Wd [H,F];  fake-quantised bf16 copies of the same shapes feed the GEMMs
  params    one flat fp32 buffer  [LET scales | LET shifts | LWC up/low per weight]  (+ flat grad, m, v)
  acts      a1,a1q [M,H]  qkv,qkvq [M,3H]  S/dP fp32 [h,T,T] (scratch)  P [b,h,T,T]  O,Oq,h1,a2,a2q [M,H]
            gu [M,2F]  m,mq [M,F]  out [M,H]   and the matching gradient buffers
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import ops

BF16, F32 = torch.bfloat16, torch.float32


@dataclass
class QuantConfig:
    """The quantiser parameter dicts of main.py:268-303, flattened."""
    wbits: int = 4
    abits: int = 4
    group_size: int = 0
    symmetric: bool = False
    lwc: bool = True
    let: bool = True
    recip_div: bool = True   # ATen-CUDA semantics of `x / (2**n - 1)`; False = ATen-CPU (true division)

    @property
    def act_quant(self):
        return self.abits < 16


class _Unit:
    """One weight-quant launch: a row range of a concatenated master weight."""

    def __init__(self, name, group, r0, r1, K, col=None, row=None, row_op=0, shift=None):
        self.name, self.group, self.r0, self.r1, self.K = name, group, r0, r1, K
        self.col, self.row, self.row_op, self.shift = col, row, row_op, shift
        self.N = r1 - r0


class BlockEngine:
    LET_SCALES = ("qkv_s", "out_s", "fc1_s", "qkt_s")
    LET_SHIFTS = ("qkv_d", "out_d", "fc1_d")

    def __init__(self, family: str, weights: Dict[str, torch.Tensor], *, n_heads: int, seq_len: int, qcfg: QuantConfig,
                 eps: float, device="cuda", batch: int = 1, let_init: Optional[Dict[str, torch.Tensor]] = None,
                 let_lr=5e-3, lwc_lr=1e-2, wd=0.0, world_size=1, causal=True, mask: Optional[torch.Tensor] = None,
                 overlap: bool = True, n_kv_heads: Optional[int] = None, rope_theta: float = 10000.0):
        assert family in ("llama", "opt", "falcon")
        self.rope_theta = float(rope_theta)
        self.family, self.q, self.eps, self.dev = family, qcfg, float(eps), torch.device(device)
        self.h, self.T, self.b = n_heads, seq_len, batch
        self.let_lr, self.lwc_lr, self.wd, self.world = let_lr, lwc_lr, wd, world_size
        self.causal = causal and mask is None
        self.mask = None if mask is None else mask.to(self.dev, F32).contiguous()
        f = lambda t: t.detach().to(self.dev, F32).contiguous()
        falcon = family == "falcon"
        H = weights["qkv" if falcon else "q"].shape[1]
        self.H, self.d = H, H // n_heads
        # grouped-query attention (Llama-2-70B, Falcon): n_kv_heads K/V heads, each shared by `rep` Q heads (repeat_kv,
        # models/int_llama_layer.py:138-139; models/int_falcon_layer.py:76-86); Hkv = n_kv_heads * d
        self.hkv = n_kv_heads or n_heads
        assert n_heads % self.hkv == 0
        self.rep = n_heads // self.hkv
        self.Hkv = self.hkv * self.d
        self.QW = H + 2 * self.Hkv
        if falcon and qcfg.let:
            raise ValueError("falcon not yet support let")          # models/int_falcon_layer.py:329-331,:344-346
        if self.rep > 1 and (qcfg.let or family == "opt"):
            raise NotImplementedError("grouped-query attention is supported without LET only (the reference's "
                                      "smooth_q_k / smooth_fc_fc do not support it either)")
        self.M = batch * seq_len
        glu = family == "llama"
        self.glu = glu
        # ---- fp32 masters, concatenated by consumer
        self.W = {}
        if falcon:      # fused QKV, per K/V group [rep query heads | k | v] (models/int_falcon_layer.py:76-86)
            self.W["qkv"], self.W["o"] = f(weights["qkv"]), f(weights["dense"])
            assert self.W["qkv"].shape[0] == self.QW
        else:
            assert weights["k"].shape[0] == self.Hkv and weights["v"].shape[0] == self.Hkv
            self.W["qkv"] = torch.cat([f(weights["q"]), f(weights["k"]), f(weights["v"])], 0)
            self.W["o"] = f(weights["o"])
        if glu:
            self.W["fc1"] = torch.cat([f(weights["gate"]), f(weights["up"])], 0)
            self.W["fc2"] = f(weights["down"])
            self.F = weights["gate"].shape[0]
        else:
            self.W["fc1"] = f(weights["fc1"])
            self.W["fc2"] = f(weights["fc2"])
            self.F = weights["fc1"].shape[0]
        F_ = self.F
        self.norm_kind = 0 if family == "llama" else 1
        self.ln = {"ln1w": f(weights["ln1"]), "ln2w": f(weights["ln2"]),
                   "ln1b": f(weights["ln1.bias"]) if "ln1.bias" in weights else None,
                   "ln2b": f(weights["ln2.bias"]) if "ln2.bias" in weights else None}
        # raw linear biases (OPT, Falcon with bias=True), concatenated like the weights
        def rawb(names):
            if all(f"{n}.bias" in weights for n in names):
                return torch.cat([f(weights[f"{n}.bias"]) for n in names], 0)
            return None
        self.rawb = {"qkv": rawb(["qkv"] if falcon else ["q", "k", "v"]), "o": rawb(["dense"] if falcon else ["o"]),
                     "fc1": rawb(["gate", "up"]) if glu else rawb(["fc1"]), "fc2": rawb(["down"] if glu else ["fc2"])}
        let = qcfg.let
        if falcon:
            U = [_Unit("qkv", "qkv", 0, self.QW, H), _Unit("o", "o", 0, H, H),
                 _Unit("fc1", "fc1", 0, self.W["fc1"].shape[0], H), _Unit("fc2", "fc2", 0, H, F_)]
        else:
            U = [
                _Unit("q", "qkv", 0, H, H, "qkv_s", "qkt_s", 2, "qkv_d"),
                _Unit("k", "qkv", H, H + self.Hkv, H, "qkv_s", "qkt_s", 1, "qkv_d"),
                _Unit("v", "qkv", H + self.Hkv, self.QW, H, "qkv_s", "out_s", 2, "qkv_d"),
                _Unit("o", "o", 0, H, H, "out_s", None, 0, "out_d"),
                _Unit("fc1", "fc1", 0, self.W["fc1"].shape[0], H, "fc1_s", None, 0, "fc1_d"),
                _Unit("fc2", "fc2", 0, H, F_, None, None, 0, None),
            ]
        if not let:
            for u in U:
                u.col = u.row = u.shift = None
                u.row_op = 0
        self.units = U
        # ---- flat parameter buffer: [LET scales | LET shifts | LWC]
        G = lambda K: (K + qcfg.group_size - 1) // qcfg.group_size if qcfg.group_size else 1
        layout, off = {}, 0
        if let:
            for n in self.LET_SCALES + self.LET_SHIFTS:
                layout[n] = (off, H)
                off += H
        self.n_let = off
        self.group_units = {g: [u for u in U if u.group == g] for g in dict.fromkeys(u.group for u in U)}
        self.glayout = {}          # weight group -> (offset of its ups, offset of its lows, items): one launch can serve a group
        if qcfg.lwc:
            for g, us in self.group_units.items():
                tot = sum(u.N * G(u.K) for u in us)
                self.glayout[g] = (off, off + tot, tot)
                o2 = off
                for u in us:
                    n = u.N * G(u.K)
                    layout[f"{u.name}.up"] = (o2, n)
                    layout[f"{u.name}.low"] = (o2 + tot, n)
                    o2 += n
                off += 2 * tot
        self.n_params = off
        self.layout = layout
        z = lambda n, dt=F32: torch.zeros(n, device=self.dev, dtype=dt)
        self.params, self.adam_m, self.adam_v = z(max(off, 1)), z(max(off, 1)), z(max(off, 1))
        # gradients and the step's loss share one buffer: a single all-reduce per optimiser step
        self._gradbuf = z(max(off, 1) + 1)
        self.grads = self._gradbuf[:max(off, 1)]
        self.step_dev = torch.zeros(1, device=self.dev, dtype=torch.int32)
        self.p = {n: self.params[o:o + c] for n, (o, c) in layout.items()}
        self.g = {n: self.grads[o:o + c] for n, (o, c) in layout.items()}
        # group-level views of the LWC factors / their gradients: (up, low) over all units of a weight group
        self.pg = {g: (self.params[a:a + n], self.params[b:b + n]) for g, (a, b, n) in self.glayout.items()}
        self.gg = {g: (self.grads[a:a + n], self.grads[b:b + n]) for g, (a, b, n) in self.glayout.items()}
        if qcfg.lwc:
            for u in U:
                self.p[f"{u.name}.up"].fill_(4.0)     # quantizer.py:62
                self.p[f"{u.name}.low"].fill_(4.0)
        if let:
            for n in self.LET_SCALES:
                self.p[n].fill_(1.0)
            if let_init:
                for n, v in let_init.items():
                    self.p[n].copy_(v.to(self.dev, F32))
        # ---- per-unit quantiser state + bf16 operands
        self.Wq = {g: torch.empty_like(w, dtype=BF16) for g, w in self.W.items()}
        # W4A4 / W6A6 with per-channel weights: the four forward linears contract the integer codes of both operands
        # exactly (oq_gemm_s8, tcgen05 kind::i8) and apply the scales in the epilogue; the bf16 copies stay for dgrad / wgrad
        self.s8 = (qcfg.act_quant and qcfg.wbits <= 8 and qcfg.abits <= 8 and not qcfg.group_size and H % 16 == 0
                   and F_ % 16 == 0 and max(H, F_) <= 16384 and (glu or F_ <= 12288) and self.dev.type == "cuda"
                   and os.environ.get("OQ_S8", "1") != "0")
        self.W8 = {g: ops.S8Operand(w.shape[0], w.shape[1], self.dev) for g, w in self.W.items()} if self.s8 else None
        self.Wfp = None
        F1 = self.W["fc1"].shape[0]
        self.F1 = F1
        self._mvbuf = z(4 * H + F1)         # packed W@shift outputs (layout: include/omniquant_b200.h, oq_let_fold_vec)
        mvv = {"q": self._mvbuf[:H], "k": self._mvbuf[H:2 * H], "v": self._mvbuf[2 * H:3 * H], "o": self._mvbuf[3 * H:4 * H],
               "fc1": self._mvbuf[4 * H:], "fc2": z(H)}
        self.qstat, self.gstat = {}, {}
        for g, us in self.group_units.items():
            tot = sum(u.N * G(u.K) for u in us)
            gs = dict(xmin=z(tot), xmax=z(tot), zp=z(tot), scale=self.W8[g].scale if (self.s8 and tot == self.W[g].shape[0]) else z(tot))
            self.gstat[g] = gs
            o2 = 0
            for u in us:
                n = u.N * G(u.K)
                self.qstat[u.name] = dict(xmin=gs["xmin"][o2:o2 + n], xmax=gs["xmax"][o2:o2 + n], zp=gs["zp"][o2:o2 + n],
                                          scale=gs["scale"][o2:o2 + n], mv=mvv.get(u.name))
                o2 += n
        self._tmpv = z(9 * H + F1)
        self.tmp = {}     # folded vectors of the current step
        self.bufs = {}
        self.loss_slot = self._gradbuf[max(off, 1):]
        self._accbuf = z(10 * H)
        self._acc = {k: self._accbuf[i * H:(i + 1) * H] for i, k in enumerate(
            ("dcol_qkv", "dcol_out", "dcol_fc1", "dsh_qkv", "dsh_out", "dsh_fc1", "dw1", "db1", "dw2", "db2"))}
        self._drowbuf = z(3 * H)
        self._drow = {k: self._drowbuf[i * H:(i + 1) * H] for i, k in enumerate(("q", "k", "v"))}
        self._qscale = self.d ** -0.5 if family == "opt" else 1.0
        self.overlap = overlap and self.dev.type == "cuda" and os.environ.get("OQ_OVERLAP", "1") != "0"
        self.side = torch.cuda.Stream(device=self.dev) if self.overlap else None
        self._pending = False
        self._qstat_home = {n: (st["scale"], st["zp"]) for n, st in self.qstat.items()}
        self._let_init_names = tuple(let_init) if let_init else ()
        self._graph_state = None      # captured optimiser step, kept across reload(): see train()
        # QK^T -> softmax -> PV (and the recompute backward) in one tcgen05 kernel each when the shape allows it
        # (csrc/attn_sm100.cu: causal, head_dim 64 or 128, T %% 128 == 0); otherwise batched GEMMs + softmax kernels
        self.fused_attn = (self.causal and self.dev.type == "cuda" and os.environ.get("OQ_FUSED_ATTN", "1") != "0"
                           and ops.attn_supported(seq_len, self.d))

    def reload(self, weights: Dict[str, torch.Tensor], let_init: Optional[Dict[str, torch.Tensor]] = None):
        """Next decoder block of the same shape into the SAME device buffers: raw weights / norms / biases are copied in
        place, LWC / LET parameters and the AdamW state are re-initialised (quantize/omniquant.py:182-208 per block).
        Every buffer keeps its address, so the CUDA graph of the optimiser step captured for the previous block
        (engine.train) is replayed as is -- capturing costs 0.1 s, and 0.6 s with the NCCL all-reduce inside."""
        f = lambda t: t.detach().to(self.dev, F32)
        falcon, glu, H, Hkv = self.family == "falcon", self.glu, self.H, self.Hkv
        with torch.no_grad():
            if falcon:
                self.W["qkv"].copy_(f(weights["qkv"])); self.W["o"].copy_(f(weights["dense"]))
            else:
                self.W["qkv"][:H].copy_(f(weights["q"])); self.W["qkv"][H:H + Hkv].copy_(f(weights["k"]))
                self.W["qkv"][H + Hkv:].copy_(f(weights["v"])); self.W["o"].copy_(f(weights["o"]))
            if glu:
                F_ = self.F
                self.W["fc1"][:F_].copy_(f(weights["gate"])); self.W["fc1"][F_:].copy_(f(weights["up"]))
                self.W["fc2"].copy_(f(weights["down"]))
            else:
                self.W["fc1"].copy_(f(weights["fc1"])); self.W["fc2"].copy_(f(weights["fc2"]))
            for key, src in (("ln1w", "ln1"), ("ln2w", "ln2"), ("ln1b", "ln1.bias"), ("ln2b", "ln2.bias")):
                if (self.ln[key] is None) != (src not in weights):
                    raise ValueError(f"reload: block layout differs from the engine's ({src})")
                if self.ln[key] is not None:
                    self.ln[key].copy_(f(weights[src]))
            groups = {"qkv": ["qkv"] if falcon else ["q", "k", "v"], "o": ["dense"] if falcon else ["o"],
                      "fc1": ["gate", "up"] if glu else ["fc1"], "fc2": ["down"] if glu else ["fc2"]}
            for g, names in groups.items():
                have = all(f"{n}.bias" in weights for n in names)
                if (self.rawb[g] is None) == have:
                    raise ValueError(f"reload: block layout differs from the engine's (bias of {g})")
                if have:
                    self.rawb[g].copy_(torch.cat([f(weights[f"{n}.bias"]) for n in names], 0))
            if self.Wfp is not None:
                for g, w in self.W.items():
                    self.Wfp[g].copy_(w)
            self.params.zero_(); self.adam_m.zero_(); self.adam_v.zero_(); self._gradbuf.zero_(); self.step_dev.zero_()
            if self.q.lwc:
                for u in self.units:
                    self.p[f"{u.name}.up"].fill_(4.0)
                    self.p[f"{u.name}.low"].fill_(4.0)
            if self.q.let:
                for n in self.LET_SCALES:
                    self.p[n].fill_(1.0)
                for n, v in (let_init or {}).items():
                    self.p[n].copy_(v.to(self.dev, F32))
            for n, (sc, zp) in self._qstat_home.items():     # prepare(final=True) re-pointed these at its outputs
                self.qstat[n]["scale"], self.qstat[n]["zp"] = sc, zp
        self.tmp = {}
        self.saved = None
        return self

    # ------------------------------------------------------------------ side stream
    # The quantiser kernels are instruction-bound and the persistent GEMM leaves SMs idle in its last wave, so
    # the weight fake-quant of the later weights (forward) and the per-weight quantiser backward run on a second
    # stream beside the GEMMs of the main stream.  fork = side waits for main; join = main waits for side.
    def _fork(self):
        if self.overlap:
            self.side.wait_stream(torch.cuda.current_stream())
            self._pending = True

    def _join(self):
        if self._pending:
            torch.cuda.current_stream().wait_stream(self.side)
            self._pending = False

    def _on_side(self):
        import contextlib
        return torch.cuda.stream(self.side) if self.overlap else contextlib.nullcontext()

    # ------------------------------------------------------------------ helpers
    def buf(self, name, shape, dtype=BF16):
        t = self.bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, device=self.dev, dtype=dtype)
            self.bufs[name] = t
        return t

    def wview(self, u: _Unit, store):
        return store[u.group][u.r0:u.r1]

    def lwc(self, u: _Unit):
        if not self.q.lwc:
            return None, None
        return self.p[f"{u.name}.up"], self.p[f"{u.name}.low"]

    # ------------------------------------------------------------------ step preparation (smooth_and_quant_temporary)
    def prepare(self, final: bool = False):
        """LET folds + weight fake-quant for this step (models/int_llama_layer.py:279-307 /
        models/int_opt_layer.py:385-413).  With ``final`` also returns fp32 fake-quantised weights,
        scales and zero points (smooth_and_quant_inplace + register_scales_and_zeros)."""
        q, p = self.q, self.p
        if q.let:
            ops.truncate_(self.params[:4 * self.H], 1e-2)   # the four *_smooth_scale vectors are contiguous
        extra = {}

        def quant_unit(u):
            up, low = self.lwc(u)
            st = self.qstat[u.name]
            kw = dict(scale_out=st["scale"], zp_out=st["zp"])
            if final and q.wbits >= 16:
                # 16-bit weights: LET fold only, the quantiser is the identity and leaves scale / zero point None
                # (quantizer.py:109; register_scales_and_zeros then registers None buffers, :149-153)
                extra[u.name] = dict(weight=torch.empty(u.N, u.K, device=self.dev), scales=None, zeros=None)
                kw = dict(wq_f32=extra[u.name]["weight"])
            elif final:
                G = st["xmin"].numel()
                extra[u.name] = dict(weight=torch.empty(u.N, u.K, device=self.dev), scales=torch.empty(G, device=self.dev),
                                     zeros=torch.empty(G, device=self.dev))
                kw = dict(wq_f32=extra[u.name]["weight"], scale_out=extra[u.name]["scales"], zp_out=extra[u.name]["zeros"])
                if self.s8:       # the scale lands in the integer operand and is copied out below
                    kw["scale_out"] = st["scale"]
                    st["zp"] = extra[u.name]["zeros"]
                else:
                    st["scale"], st["zp"] = extra[u.name]["scales"], extra[u.name]["zeros"]
            if self.s8 and q.wbits < 16:
                kw["s8"] = self.W8[u.group].rows(u.r0, u.r1)
            ops.weight_quant_fwd(self.wview(u, self.W), n_bits=q.wbits, symmetric=q.symmetric, group_size=q.group_size,
                                 up=up, low=low, col_scale=p.get(u.col) if u.col else None,
                                 row_scale=p.get(u.row) if u.row else None, row_op=u.row_op,
                                 shift=p.get(u.shift) if u.shift else None, matvec_out=st["mv"] if u.shift else None,
                                 recip_div=q.recip_div, wq_bf16=self.wview(u, self.Wq), xmin_out=st["xmin"],
                                 xmax_out=st["xmax"], **kw)
            if final and self.s8 and q.wbits < 16:
                extra[u.name]["scales"].copy_(st["scale"])

        t = self.tmp
        H = self.H
        ln, tv, F1 = self.ln, self._tmpv, self.F1
        fold = lambda part: ops.let_fold_vec(H, F1, self.params[:7 * H], ln["ln1w"], ln["ln1b"], ln["ln2w"], ln["ln2b"],
                                             self.rawb["qkv"], self.rawb["o"], self.rawb["fc1"], self._mvbuf, self._qscale,
                                             tv, part)
        first = [u for u in self.units if u.group == "qkv"]
        later = [u for u in self.units if u.group != "qkv"]
        if self._merged(first, final):      # q | k | v in ONE launch (row segments carry the per-weight row folds)
            g0 = first[0].group
            gs = self.gstat[g0]
            kw = dict(s8=self.W8[g0]) if (self.s8 and q.wbits < 16) else {}
            up, low = self.pg[g0] if q.lwc else (None, None)
            u0 = first[0]
            ops.weight_quant_fwd(self.W[g0], n_bits=q.wbits, symmetric=q.symmetric, group_size=q.group_size, up=up, low=low,
                                 col_scale=p.get(u0.col) if u0.col else None, shift=p.get(u0.shift) if u0.shift else None,
                                 matvec_out=self._mvbuf[:self.QW] if u0.shift else None, recip_div=q.recip_div,
                                 wq_bf16=self.Wq[g0], xmin_out=gs["xmin"], xmax_out=gs["xmax"], scale_out=gs["scale"],
                                 zp_out=gs["zp"], segments=self._segments(first), **kw)
        else:
            for u in first:             # what the first GEMM of the forward needs, on the main stream
                quant_unit(u)
        if q.let:
            fold(1)
        self._fork()
        with self._on_side():           # o / fc1 / fc2 beside the qkv GEMM and the attention of the main stream
            for u in later:
                quant_unit(u)
            if q.let:
                fold(2)
        if final:
            self._join()
        if q.let:
            t.update(ln1w=tv[:H], ln1b=tv[H:2 * H], ln2w=tv[2 * H:3 * H], ln2b=tv[3 * H:4 * H], bqkv=tv[4 * H:7 * H],
                     bo=tv[7 * H:8 * H], bfc1=tv[8 * H:8 * H + F1], bfc2=self.rawb["fc2"])
            if self.family == "opt":
                t["bq_scaled"] = tv[8 * H + F1:9 * H + F1]
        else:
            t.update(ln1w=self.ln["ln1w"], ln1b=self.ln["ln1b"], ln2w=self.ln["ln2w"], ln2b=self.ln["ln2b"],
                     bqkv=self.rawb["qkv"], bo=self.rawb["o"], bfc1=self.rawb["fc1"], bfc2=self.rawb["fc2"])
        if self.family == "opt" and not q.let:
            # q = (x W^T + b) * d^-0.5 is folded into the q GEMM epilogue (models/int_opt_layer.py:96)
            t["bq_scaled"] = None if t["bqkv"] is None else t["bqkv"][:H] * self.d ** -0.5
        t["W"] = self.Wq
        t["W8"] = self.W8
        return extra

    def _merged(self, us, final=False):
        """One weight-quantiser launch for all units of a weight group (q | k | v): always possible without LET (no
        fold: the units differ in nothing); with LET when the per-channel streaming kernels apply (row segment table)."""
        q = self.q
        if final or len(us) < 2 or q.wbits >= 16 or os.environ.get("OQ_WQ_MERGE", "1") == "0":
            return False
        if not q.let:
            return True
        u0 = us[0]
        return (not q.group_size and u0.K % 8 == 0 and u0.K <= 8192 and all(u.col == u0.col and u.shift == u0.shift for u in us))

    def _segments(self, us):
        if not self.q.let:
            return None
        return [(u.r1, self.p.get(u.row) if u.row else None, u.row_op if u.row else 0) for u in us]

    def fp_weights(self):
        """bf16 copies of the raw weights for the full-precision pass (quantize/omniquant.py:165-172)."""
        if self.Wfp is None:
            self.Wfp = {g: w.to(BF16) for g, w in self.W.items()}
        fp = dict(W=self.Wfp, ln1w=self.ln["ln1w"], ln1b=self.ln["ln1b"], ln2w=self.ln["ln2w"], ln2b=self.ln["ln2b"],
                  bqkv=self.rawb["qkv"], bo=self.rawb["o"], bfc1=self.rawb["fc1"], bfc2=self.rawb["fc2"])
        if self.family == "opt":
            fp["bq_scaled"] = None if fp["bqkv"] is None else fp["bqkv"][:self.H] * self.d ** -0.5
        return fp

    # ------------------------------------------------------------------ forward
    def _s8buf(self, name, shape):
        """Integer twin of an activation operand (None when the exact-integer forward is off for this pass)."""
        if not self._use_s8:
            return None
        key = name + ".s8"
        t = self.bufs.get(key)
        if t is None or tuple(t.codes.shape) != tuple(shape):
            t = self.bufs[key] = ops.S8Operand(shape[0], shape[1], self.dev)
        return t

    def _aq(self, x, name):
        """per-token fake-quant over the last dim (QuantLinear.act_quantizer); returns (bf16 operand, integer operand)."""
        if not self._act_quant:
            return x, None
        y = self.buf(name, x.shape)
        x8 = self._s8buf(name, x.shape)
        ops.act_quant_fwd(x, y, n_bits=self.q.abits, recip_div=self.q.recip_div, s8=x8)
        return y, x8

    def _norm_aq(self, x, w, b, y, name, stats):
        """y = norm(x) and, when activations are quantised, its per-token fake-quant in the same pass."""
        if not self._act_quant:
            ops.norm_fwd(x, w, b, self.eps, self.norm_kind, y, *stats)
            return y, None
        yq = self.buf(name, y.shape)
        y8 = self._s8buf(name, y.shape)
        ops.norm_act_quant_fwd(x, w, b, self.eps, self.norm_kind, y, yq, n_bits=self.q.abits, recip_div=self.q.recip_div,
                               mean_out=stats[0], rstd_out=stats[1], s8=y8)
        return yq, y8

    def _linear(self, xq, x8, ws, g, out, bias=None, residual=None, rows=None, alpha=1.0):
        """QuantLinear forward GEMM of weight group g (optionally a row range of it): exact-integer contraction when both
        operands carry their codes, bf16 tcgen05 GEMM otherwise (quantize/int_linear.py:62)."""
        r0, r1 = rows if rows is not None else (0, ws["W"][g].shape[0])
        if x8 is not None:
            w8 = ws["W8"][g]
            ops.linear_s8_fwd(x8, w8 if rows is None else w8.rows(r0, r1), out, bias=bias, residual=residual, alpha=alpha)
            return
        Wg = ws["W"][g][r0:r1]
        M, K = xq.shape
        ops.gemm(xq, Wg, out, M=M, N=r1 - r0, K=K, lda=xq.stride(0), ldb=Wg.stride(0), ldc=out.stride(0), bias=bias,
                 residual=residual, ldr=residual.stride(0) if residual is not None else 0, alpha=alpha)

    def forward(self, x: torch.Tensor, ws: dict, act_quant: bool, out: Optional[torch.Tensor] = None):
        """Block forward on x [M,H] bf16 (models/int_llama_layer.py:213-267 / int_opt_layer.py:268-346)."""
        M, H, F_, h, d, T = x.shape[0], self.H, self.F, self.h, self.d, self.T
        nb = M // T
        self._act_quant = act_quant and self.q.act_quant
        self._use_s8 = self._act_quant and self.s8 and ws.get("W8") is not None
        B = self.buf
        self.x = x
        falcon = self.family == "falcon"
        a1 = B("a1", (M, H))
        self.st1 = (B("mean1", (M,), F32), B("rstd1", (M,), F32))
        a1q, a1q8 = self._norm_aq(x, ws["ln1w"], ws["ln1b"], a1, "a1q", self.st1)
        QW, Hkv, hkv, rep = self.QW, self.Hkv, self.hkv, self.rep
        ok, ov = H, H + Hkv                     # column offsets of k and v in the [M, QW] qkv buffers (LLaMA / OPT)
        qkv = B("qkv", (M, QW))
        W = ws["W"]
        if self.family in ("llama", "falcon"):
            self._linear(a1q, a1q8, ws, "qkv", qkv, bias=ws["bqkv"])
        else:
            self._linear(a1q, a1q8, ws, "qkv", qkv[:, :H], bias=ws["bq_scaled"], rows=(0, H), alpha=d ** -0.5)
            self._linear(a1q, a1q8, ws, "qkv", qkv[:, H:], bias=None if ws["bqkv"] is None else ws["bqkv"][H:], rows=(H, 3 * H))
        # ---- Q/K/V operands of QuantMatMul
        abits = self.q.abits if self._act_quant else 16
        parts = ((0, H, h), (ok, ov, hkv), (ov, QW, hkv))     # (column range, heads) of q, k, v
        if self.family == "llama":
            qkvq = B("qkvq", (M, QW))
            cos, sin = self.rope()
            # one launch over the [q heads | k heads | v heads] of every token: rotary on q and k, per-(token, head)
            # quantisation of all three (int_llama_layer.py:124-125,141-142,162); abits >= 16: rotary only, v copied
            nh3 = h + 2 * hkv
            ops.act_quant_fwd(qkv.view(M, nh3, d), qkvq.view(M, nh3, d), n_bits=abits, recip_div=self.q.recip_div,
                              cos=cos, sin=sin, seq_len=T, rope_inner=h + hkv)
            alpha = 1.0 / math.sqrt(d)
        elif falcon:
            # rotary on the q and k heads of every group, v untouched; no Q/K/V quantisation in the Falcon block
            # (plain scaled_dot_product_attention, models/int_falcon_layer.py:148,170-176)
            qkvq = B("qkvq", (M, QW))
            cos, sin = self.rope()
            ops.act_quant_fwd(qkv.view(M, hkv * (rep + 2), d), qkvq.view(M, hkv * (rep + 2), d), n_bits=16, cos=cos, sin=sin,
                              seq_len=T)
            qkvq.view(M, hkv, rep + 2, d)[:, :, -1].copy_(qkv.view(M, hkv, rep + 2, d)[:, :, -1])
            alpha = 1.0 / math.sqrt(d)
        else:
            if abits < 16:          # per token over the full hidden dim, before the head split (int_opt_layer.py:97,125,129)
                qkvq = B("qkvq", (M, QW))
                for c0, c1, _ in parts:
                    ops.act_quant_fwd(qkv[:, c0:c1], qkvq[:, c0:c1], n_bits=abits, recip_div=self.q.recip_div)
            else:
                qkvq = qkv
            alpha = 1.0
        self.qkv, self.qkvq, self.att_alpha = qkv, qkvq, alpha
        O = B("O", (M, H))
        if self.fused_attn:
            P = None
            self.att_stats = (B("att_max", (nb, h, T), F32), B("att_invl", (nb, h, T), F32))
            ops.attn_fwd(qkvq, O, *self.att_stats, batch=nb, T=T, heads=h, kv_heads=hkv, head_dim=d, layout=self._att_cols(),
                         alpha=alpha)
        else:
            P = B("P", (nb, h, T, T))
            passes = self._att_passes()
            hmax = max(c1 - c0 for c0, c1, _ in passes)
            S = B("S", (hmax, T, T), torch.float16)     # scores in fp16 like the reference's autocast (int_llama_layer.py:143)
        cz = 1 if self.causal else 0
        for s in range(0 if self.fused_attn else nb):
            r = slice(s * T, (s + 1) * T)
            for c0, c1, groups in passes:       # a few heads at a time: scores + probabilities of a pass stay in L2
                for h0, nh, cq, ck, cv, sb in groups:
                    ops.gemm(qkvq[r, cq:], qkvq[r, ck:], S[h0 - c0:], M=T, N=T, K=d, lda=QW, ldb=QW, ldc=T, batch=nh, strideA=d,
                             strideB=sb, strideC=T * T, alpha=alpha, causal=cz)
                ops.softmax_fwd(S[:c1 - c0], P[s, c0:c1], mask=self.mask, causal=self.causal)
                for h0, nh, cq, ck, cv, sb in groups:
                    ops.gemm(P[s, h0:], qkvq[r, cv:], O[r, h0 * d:], M=T, N=d, K=T, lda=T, ldb=QW, b_mn=True, ldc=H, batch=nh,
                             strideA=T * T, strideB=sb, strideC=d, causal=2 * cz)
        Oq, Oq8 = self._aq(O, "Oq")
        h1 = B("h1", (M, H))
        self._join()                # o / fc1 / fc2 operands of this step come from the side stream
        self._linear(Oq, Oq8, ws, "o", h1, bias=ws["bo"], residual=x)
        a2 = B("a2", (M, H))
        self.st2 = (B("mean2", (M,), F32), B("rstd2", (M,), F32))
        # Falcon: ln_mlp acts on the block input, attention and MLP run in parallel (int_falcon_layer.py:272-301)
        a2q, a2q8 = self._norm_aq(x if falcon else h1, ws["ln2w"], ws["ln2b"], a2, "a2q", self.st2)
        f1 = B("f1", (M, W["fc1"].shape[0]))
        self._linear(a2q, a2q8, ws, "fc1", f1, bias=ws["bfc1"])
        if self.glu and self._act_quant:
            # SwiGLU and down_proj's input quantiser in one pass; m itself is re-derived in the backward
            m = None
            mq = B("mq", (M, F_))
            mq8 = self._s8buf("mq", (M, F_))
            ops.swiglu_act_quant_fwd(f1, mq, n_bits=self.q.abits, recip_div=self.q.recip_div, s8=mq8)
        else:
            m = B("m", (M, F_))
            if self.glu:
                ops.swiglu_fwd(f1, m)
            elif falcon:
                ops.gelu_fwd(f1, m)
            else:
                ops.relu_fwd(f1, m)
            mq, mq8 = self._aq(m, "mq")
        if out is None:
            out = B("out", (M, H))
        self._linear(mq, mq8, ws, "fc2", out, bias=ws["bfc2"], residual=h1)
        self.saved = dict(a1=a1, a1q=a1q, O=O, Oq=Oq, h1=h1, a2=a2, a2q=a2q, f1=f1, m=m, mq=mq, P=P, ws=ws)
        return out

    def _att_cols(self):
        """Column layout of the heads inside the [M, QW] qkv buffers for the fused attention kernels
        (include/omniquant_b200.h, oq_attn_fwd): LLaMA / OPT store [q | k | v], Falcon per K/V group [rep q heads | k | v]."""
        d, rep = self.d, self.rep
        if self.family == "falcon":
            gw = (rep + 2) * d
            return dict(cq=0, gq=gw, ck=rep * d, cv=(rep + 1) * d, gk=gw)
        return dict(cq=0, gq=rep * d, ck=self.H, cv=self.H + self.Hkv, gk=d)

    def _att_layout(self):
        """Launch groups of the attention GEMMs: (first head, heads in the launch, q / k / v column offsets in the
        [M, QW] qkv buffers, batch stride of the K/V operand -- 0 = shared by the launch's heads).
        LLaMA / OPT store [q | k | v]; Falcon stores per K/V group [rep q heads | k | v]."""
        h, d, H, hkv, rep, Hkv = self.h, self.d, self.H, self.hkv, self.rep, self.Hkv
        if self.family == "falcon":
            gw = (rep + 2) * d
            return [(g * rep, rep, g * gw, g * gw + rep * d, g * gw + (rep + 1) * d, 0 if rep > 1 else d) for g in range(hkv)]
        if rep == 1:
            return [(0, h, 0, H, H + Hkv, d)]
        return [(g * rep, rep, g * rep * d, H + g * d, H + Hkv + g * d, 0) for g in range(hkv)]

    def _att_passes(self):
        """[(first head, end head, launch groups)]: optionally the attention runs over a few heads at a time so that the
        fp16 scores and bf16 probabilities of one pass (causal: half of heads*T*T*4 bytes) fit the 126 MB L2 between the
        QK^T GEMM, the softmax and the P.V GEMM.  OQ_ATT_PASS_MB sets the budget; default 0 = all heads in one pass:
        measured on B200 at LLaMA-7B shapes the split is SLOWER (3.49 ms/step single pass, 3.56 ms at 72 MB, 3.81 ms at
        40 MB) -- the extra launches and partial waves cost more than the HBM trips they save."""
        if getattr(self, "_passes", None) is not None:
            return self._passes
        import os
        budget = float(os.environ.get("OQ_ATT_PASS_MB", "0")) * 1e6
        per_head = self.T * self.T * 4 * (0.5 if self.causal else 1.0)
        hc = self.h if budget <= 0 else max(1, min(self.h, int(budget // per_head)))
        groups = self._att_layout()
        if len(groups) == 1 and groups[0][1] == self.h and self.rep == 1:      # MHA: one launch group, split by heads
            h0, nh, cq, ck, cv, sb = groups[0]
            d = self.d
            n_pass = -(-self.h // hc)
            hc = -(-self.h // n_pass)
            groups = [(a, min(hc, self.h - a), cq + a * d, ck + a * d, cv + a * d, sb) for a in range(0, self.h, hc)]
        passes, cur = [], []
        for g in groups:                          # whole launch groups (K/V groups under GQA) per pass
            if cur and (g[0] + g[1] - cur[0][0]) > hc:
                passes.append((cur[0][0], cur[-1][0] + cur[-1][1], cur))
                cur = []
            cur.append(g)
        passes.append((cur[0][0], cur[-1][0] + cur[-1][1], cur))
        self._passes = passes
        return passes

    def rope(self):
        if not hasattr(self, "_rope"):
            # LlamaRotaryEmbedding (transformers 4.31, deep-copied by the reference at int_llama_layer.py:70): base from
            # config.rope_theta; scaled variants (linear / NTK / yarn) are rejected by the block wrappers
            inv = 1.0 / (self.rope_theta ** (torch.arange(0, self.d, 2, dtype=F32) / self.d))
            emb = torch.outer(torch.arange(self.T, dtype=F32), inv)
            emb = torch.cat((emb, emb), -1)
            self._rope = (emb.cos().to(self.dev).contiguous(), emb.sin().to(self.dev).contiguous())
        return self._rope

    # ------------------------------------------------------------------ backward
    def _aq_bwd(self, x, gy, name):
        if not self._act_quant:
            return gy
        gx = self.buf(name, x.shape)
        ops.act_quant_bwd(x, gy, gx, n_bits=self.q.abits, recip_div=self.q.recip_div)
        return gx

    def backward(self, dout: torch.Tensor):
        """Hand-derived backward of forward() down to the LWC / LET gradients."""
        sv, q, B = self.saved, self.q, self.buf
        ws, W = sv["ws"], sv["ws"]["W"]
        M, H, F_, h, d, T = dout.shape[0], self.H, self.F, self.h, self.d, self.T
        nb = M // T
        let = q.let
        gdt = BF16
        dW = {g: B(f"dW_{g}", w.shape, gdt) for g, w in self.W.items()}
        if let:
            self._accbuf.zero_()
        # fc2 / down_proj
        dmq = B("dmq", (M, F_))
        ops.linear_dgrad(dout, W["fc2"], dmq)
        ops.linear_wgrad(dout, sv["mq"], dW["fc2"])
        self._unit_backward_side("fc2", dW, None)
        df1 = B("df1", sv["f1"].shape)
        falcon = self.family == "falcon"
        fused_glu = self.glu and self._act_quant
        dm = None if fused_glu else self._aq_bwd(sv["m"], dmq, "dm")
        if fused_glu:
            ops.swiglu_act_quant_bwd(sv["f1"], dmq, df1, n_bits=q.abits, recip_div=q.recip_div)
        elif self.glu:
            ops.swiglu_bwd(sv["f1"], dm, df1)
        elif falcon:
            ops.gelu_bwd(sv["f1"], dm, df1)
        else:
            ops.relu_bwd(sv["f1"], dm, df1)
        ops.linear_wgrad(df1, sv["a2q"], dW["fc1"])
        self._unit_backward_side("fc1", dW, df1)
        if falcon:
            dh1 = dout      # parallel block, LWC only: the MLP branch does not feed the attention branch
        else:
            da2q = B("da2q", (M, H))
            ops.linear_dgrad(df1, W["fc1"], da2q)
            da2 = self._aq_bwd(sv["a2"], da2q, "da2")
            dh1 = B("dh1", (M, H))
            ops.norm_bwd(sv["h1"], ws["ln2w"], self.norm_kind, self.st2[0], self.st2[1], da2, dx=dh1, dx_add=dout,
                         dw_acc=self._acc["dw2"] if let else None, db_acc=self._acc["db2"] if let else None)
        # out projection
        dOq = B("dOq", (M, H))
        ops.linear_dgrad(dh1, W["o"], dOq)
        ops.linear_wgrad(dh1, sv["Oq"], dW["o"])
        self._unit_backward_side("o", dW, dh1)
        dO = self._aq_bwd(sv["O"], dOq, "dO")
        # attention
        qkv, qkvq, alpha = self.qkv, self.qkvq, self.att_alpha
        QW, Hkv, hkv, rep = self.QW, self.Hkv, self.hkv, self.rep
        ok, ov = H, H + Hkv
        dqkvq = B("dqkvq", (M, QW))
        cz = 1 if self.causal else 0
        shared = rep > 1
        if self.fused_attn:
            ops.attn_bwd(qkvq, sv["O"], dO, *self.att_stats, B("att_delta", (nb, h, T), F32), dqkvq, batch=nb, T=T, heads=h,
                         kv_heads=hkv, head_dim=d, layout=self._att_cols(), alpha=alpha)
        else:
            passes = self._att_passes()
            hmax = max(c1 - c0 for c0, c1, _ in passes)
            dP = B("S", (hmax, T, T), torch.float16).view(BF16)   # same scratch storage as the scores
            dkv32 = B("dkv32", (T, 2 * Hkv), F32) if shared else None     # K/V gradients summed over the group's Q heads
        for s in range(0 if self.fused_attn else nb):
            r = slice(s * T, (s + 1) * T)
            P = sv["P"][s]
            for c0, c1, groups in passes:
                for h0, nh, cq, ck, cv, sb in groups:
                    ops.gemm(dO[r, h0 * d:], qkvq[r, cv:], dP[h0 - c0:], M=T, N=T, K=d, lda=H, ldb=QW, ldc=T, batch=nh, strideA=d,
                             strideB=sb, strideC=T * T, causal=cz)
                if not shared:
                    for h0, nh, cq, ck, cv, sb in groups:
                        ops.gemm(P[h0:], dO[r, h0 * d:], dqkvq[r, cv:], M=T, N=d, K=T, a_mn=True, lda=T, b_mn=True, ldb=H, ldc=QW,
                                 batch=nh, strideA=T * T, strideB=d, strideC=d, causal=3 * cz)
                else:
                    for hh in range(c0, c1):     # dV_g = sum over the group's heads of P_h^T dO_h
                        kv0 = (hh // rep) * d
                        ops.gemm(P[hh], dO[r, hh * d:], dkv32[:, Hkv + kv0:], M=T, N=d, K=T, a_mn=True, lda=T, b_mn=True,
                                 ldb=H, ldc=2 * Hkv, causal=3 * cz, accumulate=(hh % rep != 0))
                ops.softmax_bwd(P[c0:c1], dP[:c1 - c0], P[c0:c1], causal=self.causal)      # dS overwrites P
                for h0, nh, cq, ck, cv, sb in groups:
                    ops.gemm(P[h0:], qkvq[r, ck:], dqkvq[r, cq:], M=T, N=d, K=T, lda=T, b_mn=True, ldb=QW, ldc=QW, batch=nh,
                             strideA=T * T, strideB=sb, strideC=d, alpha=alpha, causal=2 * cz)
                if not shared:
                    for h0, nh, cq, ck, cv, sb in groups:
                        ops.gemm(P[h0:], qkvq[r, cq:], dqkvq[r, ck:], M=T, N=d, K=T, a_mn=True, lda=T, b_mn=True, ldb=QW, ldc=QW,
                                 batch=nh, strideA=T * T, strideB=d, strideC=d, alpha=alpha, causal=3 * cz)
                else:
                    allg = self._att_layout()
                    for hh in range(c0, c1):     # dK_g = sum over the group's heads of dS_h^T Q_h
                        kv0 = (hh // rep) * d
                        cq = allg[hh // rep][2] + (hh % rep) * d
                        ops.gemm(P[hh], qkvq[r, cq:], dkv32[:, kv0:], M=T, N=d, K=T, a_mn=True, lda=T, b_mn=True, ldb=QW,
                                 ldc=2 * Hkv, alpha=alpha, causal=3 * cz, accumulate=(hh % rep != 0))
            if shared:
                if falcon:
                    gv = dqkvq[r].view(T, hkv, rep + 2, d)
                    gv[:, :, -2].copy_(dkv32[:, :Hkv].view(T, hkv, d))
                    gv[:, :, -1].copy_(dkv32[:, Hkv:].view(T, hkv, d))
                else:
                    dqkvq[r, ok:].copy_(dkv32)
        abits = q.abits if self._act_quant else 16
        parts = ((0, H, h), (ok, ov, hkv), (ov, QW, hkv))
        if self.family == "llama":
            dqkv = B("dqkv", (M, QW))
            cos, sin = self.rope()
            nh3 = h + 2 * hkv
            ops.act_quant_bwd(qkv.view(M, nh3, d), dqkvq.view(M, nh3, d), dqkv.view(M, nh3, d), n_bits=abits,
                              recip_div=q.recip_div, cos=cos, sin=sin, seq_len=T, rope_inner=h + hkv)
        elif falcon:
            dqkv = B("dqkv", (M, QW))
            cos, sin = self.rope()
            nhh = hkv * (rep + 2)
            ops.act_quant_bwd(qkv.view(M, nhh, d), dqkvq.view(M, nhh, d), dqkv.view(M, nhh, d), n_bits=16, cos=cos, sin=sin,
                              seq_len=T)
            dqkv.view(M, hkv, rep + 2, d)[:, :, -1].copy_(dqkvq.view(M, hkv, rep + 2, d)[:, :, -1])     # v: no rotary
        else:
            if abits < 16:
                dqkv = B("dqkv", (M, QW))
                for c0, c1, _ in parts:
                    ops.act_quant_bwd(qkv[:, c0:c1], dqkvq[:, c0:c1], dqkv[:, c0:c1], n_bits=abits, recip_div=q.recip_div)
            else:
                dqkv = dqkvq
        # qkv projections
        if self.family in ("llama", "falcon"):
            ops.linear_wgrad(dqkv, sv["a1q"], dW["qkv"])
        else:
            sc = d ** -0.5   # q = sc * (x W^T + b)
            ops.gemm(dqkv, sv["a1q"], dW["qkv"], M=H, N=H, K=M, a_mn=True, lda=QW, b_mn=True, ldb=H, ldc=H, alpha=sc)
            ops.gemm(dqkv[:, H:], sv["a1q"], dW["qkv"][H:], M=2 * H, N=H, K=M, a_mn=True, lda=QW, b_mn=True, ldb=H, ldc=H)
        if let:
            da1q = B("da1q", (M, H))
            if self.family == "llama":
                ops.linear_dgrad(dqkv, W["qkv"], da1q)
            else:
                ops.gemm(dqkv[:, H:], W["qkv"][H:], da1q, M=M, N=H, K=2 * H, lda=QW, b_mn=True, ldb=H, ldc=H)
                ops.gemm(dqkv, W["qkv"], da1q, M=M, N=H, K=H, lda=QW, b_mn=True, ldb=H, ldc=H, alpha=d ** -0.5,
                         residual=da1q, ldr=H)
            da1 = self._aq_bwd(sv["a1"], da1q, "da1")
            ops.norm_bwd(self.x, ws["ln1w"], self.norm_kind, self.st1[0], self.st1[1], da1, dw_acc=self._acc["dw1"],
                         db_acc=self._acc["db1"])
        self._weight_backward(dW, dqkv, dh1, df1)

    def _unit_bwd(self, u, dW, e_row):
        """Quantiser + LET-fold backward of one weight (csrc/weight_quant*.cu)."""
        q, p, g, acc = self.q, self.p, self.g, self._acc
        if q.wbits >= 16 and not q.let:
            return          # identity quantiser, nothing folded: no parameter depends on this weight
        colacc = {"qkv_s": acc["dcol_qkv"], "out_s": acc["dcol_out"], "fc1_s": acc["dcol_fc1"]}
        shacc = {"qkv_d": acc["dsh_qkv"], "out_d": acc["dsh_out"], "fc1_d": acc["dsh_fc1"]}
        up, low = self.lwc(u)
        st = self.qstat[u.name]
        ops.weight_quant_bwd(self.wview(u, self.W), self.wview(u, dW), st["xmin"], st["xmax"], n_bits=q.wbits,
                             symmetric=q.symmetric, group_size=q.group_size, up=up, low=low,
                             col_scale=p.get(u.col) if u.col else None, row_scale=p.get(u.row) if u.row else None,
                             row_op=u.row_op, recip_div=q.recip_div, e_row=e_row if u.shift else None,
                             d_up=g.get(f"{u.name}.up"), d_low=g.get(f"{u.name}.low"),
                             dcol_acc=colacc[u.col] if u.col else None, drow=self._drow[u.name] if u.row else None,
                             dshift_acc=shacc[u.shift] if u.shift else None,
                             scale=st["scale"] if q.wbits < 16 else None, zp=st["zp"] if q.wbits < 16 else None)

    def _unit_backward_side(self, name, dW, dy):
        """Backward of the o / fc1 / fc2 weight on the side stream, as soon as its wgrad GEMM has been issued:
        it overlaps the remaining GEMMs of the main stream.  dy: the linear's output gradient (for the bias /
        shift path), None when the weight carries no LET shift."""
        u = next(x for x in self.units if x.name == name)
        self._fork()
        with self._on_side():
            e = None
            if self.q.let and u.shift and dy is not None:
                e = self.buf(f"db_{name}", (dy.shape[1],), F32)
                ops.colsum(dy, e)
            self._unit_bwd(u, dW, e)

    def _weight_backward(self, dW, dqkv, dh1, df1):
        """q / k / v weights on the main stream, then join the side stream and assemble the LET gradients."""
        q, H = self.q, self.H
        let = q.let
        e = {}
        if let:
            dbqkv = self.buf("dbqkv", (3 * H,), F32)
            ops.colsum(dqkv, dbqkv)
            ev = self.buf("e_qkv", (3 * H,), F32)
            ops.let_prep_bwd(H, self.params[:7 * H], dbqkv, self._qscale, ev)
            e = {"q": ev[:H], "k": ev[H:2 * H], "v": ev[2 * H:]}
        first = [u for u in self.units if u.group == "qkv"]
        if self._merged(first):
            g0, u0 = first[0].group, first[0]
            gs, acc = self.gstat[g0], self._acc
            colacc = {"qkv_s": acc["dcol_qkv"], "out_s": acc["dcol_out"], "fc1_s": acc["dcol_fc1"]}
            shacc = {"qkv_d": acc["dsh_qkv"], "out_d": acc["dsh_out"], "fc1_d": acc["dsh_fc1"]}
            up, low = self.pg[g0] if q.lwc else (None, None)
            d_up, d_low = self.gg[g0] if q.lwc else (None, None)
            ops.weight_quant_bwd(self.W[g0], dW[g0], gs["xmin"], gs["xmax"], n_bits=q.wbits, symmetric=q.symmetric,
                                 group_size=q.group_size, up=up, low=low, col_scale=self.p.get(u0.col) if u0.col else None,
                                 recip_div=q.recip_div, e_row=ev if (let and u0.shift) else None, d_up=d_up, d_low=d_low,
                                 dcol_acc=colacc[u0.col] if u0.col else None, drow=self._drowbuf if let else None,
                                 dshift_acc=shacc[u0.shift] if u0.shift else None, scale=gs["scale"], zp=gs["zp"],
                                 segments=self._segments(first))
        else:
            for u in first:
                self._unit_bwd(u, dW, e.get(u.name))
        self._join()
        if let:
            ops.let_grad_vec(H, self.F1, self.params[:7 * H], self._tmpv, self._accbuf, self._drowbuf, dbqkv, self._qscale,
                             self.grads[:7 * H])

    # ------------------------------------------------------------------ one optimiser step
    def loss_and_grad(self, out, target, target2=None):
        self.loss_slot.zero_()
        dout = self.buf("dout", out.shape)
        ops.mse_fwd_bwd(out, target, self.loss_slot, dout, target2=target2)
        return dout

    def train_step(self, x, target, target2=None, allreduce=None):
        """smooth_and_quant_temporary + forward + MSE + backward + AdamW  (quantize/omniquant.py:217-229)."""
        self.prepare()
        out = self.forward(x, self.tmp, act_quant=True)
        dout = self.loss_and_grad(out, target, target2)
        self.backward(dout)
        if allreduce is not None:
            allreduce(self._gradbuf)          # flat LWC+LET gradients + the loss scalar in one collective
        self.optimizer_step()
        return self.loss_slot

    def optimizer_step(self):
        ops.step_inc(self.step_dev)
        gs = 1.0 / self.world
        if self.n_let:
            ops.adamw(self.params[:self.n_let], self.grads[:self.n_let], self.adam_m[:self.n_let], self.adam_v[:self.n_let],
                      self.step_dev, lr=self.let_lr, weight_decay=self.wd, grad_scale=gs)
        if self.n_params > self.n_let:
            sl = slice(self.n_let, self.n_params)
            ops.adamw(self.params[sl], self.grads[sl], self.adam_m[sl], self.adam_v[sl], self.step_dev, lr=self.lwc_lr,
                      weight_decay=self.wd, grad_scale=gs)


# ---------------------------------------------------------------------- per-block calibration loop
ACT_STAT_SITES = ("qkv", "out", "fc1", "fc2")     # inputs of q/k/v, the attention output projection, fc1 (gate/up), fc2 (down)


def new_act_stats(eng: BlockEngine, n_samples: int):
    """Per-sample buffers {site: (absmax [n, C], mid [n, C])} filled by ``fp_pass(..., stats=...)``."""
    C = {"qkv": eng.H, "out": eng.H, "fc1": eng.H, "fc2": eng.F}
    return {k: (torch.zeros(n_samples, C[k], device=eng.dev), torch.zeros(n_samples, C[k], device=eng.dev))
            for k in ACT_STAT_SITES}


def fold_act_stats(stats, sample_index=None, nsamples: Optional[int] = None, allreduce=None):
    """{site: (act_scale [C], act_shift [C])} from the per-sample vectors, folded in GLOBAL sample order like
    generate_act_scale_shift.py:34-37,69-72.  With several ranks every rank scatters its rows into a zero
    [nsamples, C] matrix and the matrices are summed (x + 0 is exact), so all ranks fold identical inputs."""
    out = {}
    for site, (am, md) in stats.items():
        if allreduce is not None:
            idx = torch.as_tensor(sample_index, device=am.device)
            full_a = torch.zeros(nsamples, am.shape[1], device=am.device)
            full_m = torch.zeros_like(full_a)
            full_a[idx], full_m[idx] = am, md
            allreduce(full_a)
            allreduce(full_m)
            am, md = full_a, full_m
        scale, shift = torch.empty(am.shape[1], device=am.device), torch.empty(am.shape[1], device=am.device)
        ops.act_stats_fold(am.contiguous(), md.contiguous(), scale, shift, first=True)
        out[site] = (scale, shift)
    return out


def fp_pass(eng: BlockEngine, fp_inps: torch.Tensor, quant_inps: Optional[torch.Tensor] = None, aug_loss: bool = False,
            stats=None):
    """Full-precision targets (quantize/omniquant.py:165-172): fp_inps <- block(fp_inps) in place; with
    ``aug_loss`` also returns block_fp(quant_inps) (the reference's fp_inps_2).  ``stats`` (new_act_stats) additionally
    records the per-channel statistics of the linear-layer inputs of every sample -- the activations of the unquantised
    model that generate_act_scale_shift.py collects in a separate run."""
    n, T, H = fp_inps.shape
    b = eng.b
    fpw = eng.fp_weights()
    fp2 = None
    with torch.no_grad():
        if aug_loss:
            fp2 = torch.empty_like(quant_inps)
            for j in range(0, n, b):
                eng.forward(quant_inps[j:j + b].view(b * T, H), fpw, act_quant=False, out=fp2[j:j + b].view(b * T, H))
        tmp = torch.empty(b * T, H, device=eng.dev, dtype=BF16)
        for j in range(0, n, b):
            eng.forward(fp_inps[j:j + b].view(b * T, H), fpw, act_quant=False, out=tmp)
            if stats is not None:
                sv = eng.saved
                for site, act in (("qkv", sv["a1"]), ("out", sv["O"]), ("fc1", sv["a2"]), ("fc2", sv["m"])):
                    for i in range(b):
                        ops.act_stats(act[i * T:(i + 1) * T], stats[site][0][j + i], stats[site][1][j + i])
            fp_inps[j:j + b].view(b * T, H).copy_(tmp)
    return fp2


def train(eng: BlockEngine, fp_inps, quant_inps, *, epochs: int, fp_inps_2=None, allreduce=None, record_grads=False,
          log=None, use_graph: bool = False):
    """The AdamW loop (quantize/omniquant.py:211-234) on this rank's samples; eng.b samples per step.
    Losses stay on the device (no per-step host sync, unlike omniquant.py:223)."""
    n, T, H = quant_inps.shape
    b = eng.b
    spe = n // b
    steps = epochs * spe
    losses = torch.zeros(max(steps, 1), device=eng.dev)
    grads0 = None
    k = 0
    graph = None
    graph_ok = use_graph and steps > 3 and (allreduce is None or os.environ.get("OQ_GRAPH_NCCL", "1") == "1")
    if graph_ok:
        gkey = (allreduce is not None, fp_inps_2 is not None, b, T, eng.world)
        gs = eng._graph_state
        if gs is not None and gs["key"] == gkey and os.environ.get("OQ_GRAPH_REUSE", "1") == "1":
            sx, st, st2, graph = gs["sx"], gs["st"], gs["st2"], gs["graph"]     # captured for an earlier block (reload())
        else:
            sx = torch.empty(b * T, H, device=eng.dev, dtype=BF16)
            st = torch.empty_like(sx)
            st2 = torch.empty_like(sx) if fp_inps_2 is not None else None
    with torch.no_grad():
        for ep in range(epochs):
            for j in range(0, n, b):
                x = quant_inps[j:j + b].view(b * T, H)
                tgt = fp_inps[j:j + b].view(b * T, H)
                tgt2 = fp_inps_2[j:j + b].view(b * T, H) if fp_inps_2 is not None else None
                if record_grads and k == 0:
                    eng.prepare()
                    out = eng.forward(x, eng.tmp, act_quant=True)
                    eng.backward(eng.loss_and_grad(out, tgt, tgt2))
                    grads0 = eng.grads.clone()
                if graph_ok:
                    sx.copy_(x)
                    st.copy_(tgt)
                    if st2 is not None:
                        st2.copy_(tgt2)
                    if graph is None and k >= 2:      # two eager steps warm every buffer, then capture
                        graph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(graph):
                            eng.train_step(sx, st, st2, allreduce=allreduce)   # NCCL all-reduce is captured with the step
                        eng._graph_state = dict(key=gkey, sx=sx, st=st, st2=st2, graph=graph)
                    if graph is not None:
                        graph.replay()                # (capturing does not execute: the captured step runs here)
                    else:
                        eng.train_step(sx, st, st2, allreduce=allreduce)
                    loss = eng.loss_slot
                else:
                    loss = eng.train_step(x, tgt, tgt2, allreduce=allreduce)
                losses[k:k + 1].copy_(loss)
                k += 1
            if log is not None:
                log(ep, losses[ep * spe:(ep + 1) * spe])
    if allreduce is not None and eng.world > 1:
        losses = losses / eng.world
    return losses, grads0


def requant_pass(eng: BlockEngine, quant_inps: torch.Tensor):
    """quant_inps <- quantised block(quant_inps) in place (quantize/omniquant.py:242-245); expects
    eng.prepare(final=True) to have produced the folded, fake-quantised operands."""
    n, T, H = quant_inps.shape
    b = eng.b
    with torch.no_grad():
        tmp = torch.empty(b * T, H, device=eng.dev, dtype=BF16)
        for j in range(0, n, b):
            eng.forward(quant_inps[j:j + b].view(b * T, H), eng.tmp, act_quant=True, out=tmp)
            quant_inps[j:j + b].view(b * T, H).copy_(tmp)


def calibrate_block(eng: BlockEngine, fp_inps: torch.Tensor, quant_inps: torch.Tensor, *, epochs: int,
                    aug_loss: bool = False, allreduce=None, record_grads: bool = False, log=None, use_graph=False):
    """One decoder block of the reference loop (quantize/omniquant.py:157-250) on this rank's samples:
    FP pass, AdamW loop, in-place fold + quantise, re-quantised forward.  fp_inps / quant_inps
    ([n, T, H] bf16, resident in HBM) are overwritten with the block outputs like the reference's."""
    n, T, H = quant_inps.shape
    assert n % eng.b == 0 and T == eng.T and H == eng.H
    fp2 = fp_pass(eng, fp_inps, quant_inps, aug_loss)
    losses, grads0 = train(eng, fp_inps, quant_inps, epochs=epochs, fp_inps_2=fp2, allreduce=allreduce,
                           record_grads=record_grads, log=log, use_graph=use_graph)
    with torch.no_grad():
        final = eng.prepare(final=True)          # smooth_and_quant_inplace (omniquant.py:239)
    requant_pass(eng, quant_inps)
    return dict(losses=losses, grads0=grads0, final=final)
