"""Helpers to load the golden fixtures (tests/golden/*.npz, generated from the
live reference by tests/golden/make_golden.py) into oracle / engine inputs."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LLAMA_MAP = {"q": "self_attn.q_proj", "k": "self_attn.k_proj", "v": "self_attn.v_proj", "o": "self_attn.o_proj",
             "gate": "mlp.gate_proj", "up": "mlp.up_proj", "down": "mlp.down_proj"}
OPT_MAP = {"q": "self_attn.q_proj", "k": "self_attn.k_proj", "v": "self_attn.v_proj", "o": "self_attn.out_proj",
           "fc1": "fc1", "fc2": "fc2"}
FALCON_MAP = {"qkv": "self_attention.query_key_value", "dense": "self_attention.dense", "fc1": "mlp.dense_h_to_4h",
              "fc2": "mlp.dense_4h_to_h"}
LET_MAP = {"qkv_s": "qkv_smooth_scale", "qkv_d": "qkv_smooth_shift", "out_s": "out_smooth_scale",
           "out_d": "out_smooth_shift", "fc1_s": "fc1_smooth_scale", "fc1_d": "fc1_smooth_shift",
           "qkt_s": "qkt_smooth_scale"}


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def t(a):
    return torch.from_numpy(np.asarray(a)).float()


class BlockFixture:
    """One calibrated block from the reference."""

    def __init__(self, fname):
        z = load(fname)
        self.z = z
        self.family = fname.split("_")[0]
        self.lmap = {"llama": LLAMA_MAP, "opt": OPT_MAP, "falcon": FALCON_MAP}[self.family]
        g = lambda k: int(z[f"in/{k}"])
        self.H, self.F, self.heads, self.T = g("H"), g("F"), g("heads"), g("T")
        self.nsamples, self.epochs, self.batch_size = g("nsamples"), g("epochs"), g("batch_size")
        self.wbits, self.abits, self.group_size = g("wbits"), g("abits"), g("group_size") or None
        self.lwc, self.let, self.aug_loss = bool(g("lwc")), bool(g("let")), bool(g("aug_loss"))
        self.inps = t(z["in/inps"])
        self.kv_heads = int(z["in/kv_heads"]) if "in/kv_heads" in z else self.heads
        self.symmetric = bool(int(z["in/symmetric"])) if "in/symmetric" in z else False

    # raw weights in the oracle's short naming
    def weights(self):
        z, W = self.z, {}
        for s, long in self.lmap.items():
            W[s] = t(z[f"in/w/{long}.weight"])
            if f"in/w/{long}.bias" in z:
                W[f"{s}.bias"] = t(z[f"in/w/{long}.bias"])
        if self.family == "llama":
            W["ln1"] = t(z["in/w/input_layernorm.weight"])
            W["ln2"] = t(z["in/w/post_attention_layernorm.weight"])
        elif self.family == "falcon":
            W["ln1"], W["ln1.bias"] = t(z["in/w/ln_attn.weight"]), t(z["in/w/ln_attn.bias"])
            W["ln2"], W["ln2.bias"] = t(z["in/w/ln_mlp.weight"]), t(z["in/w/ln_mlp.bias"])
        else:
            W["ln1"], W["ln1.bias"] = t(z["in/w/self_attn_layer_norm.weight"]), t(z["in/w/self_attn_layer_norm.bias"])
            W["ln2"], W["ln2.bias"] = t(z["in/w/final_layer_norm.weight"]), t(z["in/w/final_layer_norm.bias"])
        return W

    def act_scales(self):
        z = self.z
        out = {}
        for k in z.files:
            if k.startswith("in/act_scale/"):
                out[k[len("in/act_scale/"):]] = t(z[k])
        return out

    def act_shifts(self):
        z = self.z
        return {k[len("in/act_shift/"):]: t(z[k]) for k in z.files if k.startswith("in/act_shift/")}

    def _params(self, prefix):
        """reference-named omni parameters -> oracle short names."""
        z, P = self.z, {}
        for s, long in self.lmap.items():
            for a, b in (("up", "upbound_factor"), ("low", "lowbound_factor")):
                k = f"{prefix}/{long}.weight_quantizer.{b}"
                if k in z:
                    P[f"{s}.{a}"] = t(z[k])
        for s, long in LET_MAP.items():
            k = f"{prefix}/{long}"
            if k in z:
                P[s] = t(z[k])
        return P

    def init_params(self):
        return self._params("init")

    def final_params(self):
        return self._params("final")

    def grads0(self):
        return self._params("grad0")

    def quantized(self):
        """in-place folded + fake-quantised buffers after calibration."""
        z, Q = self.z, {}
        for s in self.lmap:
            Q[f"{s}.w"] = t(z[f"qw/{s}.weight"])
            Q[f"{s}.b"] = t(z[f"qw/{s}.bias"]) if f"qw/{s}.bias" in z else None
            Q[f"{s}.scales"] = t(z[f"qw/{s}.scales"])
            Q[f"{s}.zeros"] = t(z[f"qw/{s}.zeros"])
        for k in ("ln1.weight", "ln1.bias", "ln2.weight", "ln2.bias"):
            kk = f"final/__buf/{k}"
            if kk in z:
                Q[k] = t(z[kk])
        return Q

    @property
    def losses(self):
        return self.z["out/losses"]

    @property
    def fp_out(self):
        return t(self.z["out/fp_out"])

    @property
    def quant_out(self):
        return t(self.z["out/quant_out"])
