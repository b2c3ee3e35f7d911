"""Pin oracle/oq_oracle.py against fixtures generated from the live reference
(tests/golden/make_golden.py).  fp32 CPU arithmetic: the quantiser part is
checked bit-exactly, block trajectories to float round-off (different op
fusion order inside autograd is the only source of difference)."""
import numpy as np
import pytest
import torch

from golden_util import BlockFixture, load, t
from oracle import oq_oracle as orc

KAT = load("quantizer_kat.npz")
W_CASES = sorted({k.split("/")[0] for k in KAT.files if k.startswith("w") or k.startswith("kat")})
A_CASES = sorted({k.split("/")[0] for k in KAT.files if k.startswith("a")})


def _beq(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32) == b.view(np.uint32) | (np.isnan(a) & np.isnan(b)),
                                                 np.ones(a.shape, bool))


def beq(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape, (a.shape, b.shape)
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))
    assert same.all(), f"{(~same).sum()} of {same.size} differ; first {a[~same][:4]} vs {b[~same][:4]}"


@pytest.mark.parametrize("case", W_CASES)
def test_weight_quantiser_bit_exact(case):
    g = lambda k: KAT[f"{case}/{k}"]
    nb, sym, grp, lwc = [int(v) for v in g("meta")]
    W = t(g("W"))
    up = t(g("up")) if f"{case}/up" in KAT else (torch.ones(orc.lwc_shape(W, grp or None), 1) * 4.0 if lwc else None)
    low = t(g("low")) if f"{case}/low" in KAT else up
    deq, s, z, codes = orc.fake_quant(W, nb, bool(sym), grp or None, up, low, return_all=True)
    beq(deq.numpy(), g("deq"))
    beq(s.numpy(), g("scale"))
    beq(z.numpy(), g("zp"))
    if f"{case}/codes" in KAT:
        beq(codes.numpy(), g("codes"))
    # numpy twin, fed the reference's own sigmoid values
    sig_up = g("sig_up") if f"{case}/sig_up" in KAT else (torch.sigmoid(up).numpy() if lwc else None)
    sig_low = g("sig_low") if f"{case}/sig_low" in KAT else (torch.sigmoid(low).numpy() if lwc else None)
    d2, s2, z2, c2 = orc.fake_quant_numpy(g("W"), nb, bool(sym), grp or None, sig_up, sig_low)
    beq(d2, g("deq"))
    beq(s2, g("scale"))
    beq(z2, g("zp"))
    if f"{case}/codes" in KAT:
        beq(c2, g("codes"))


@pytest.mark.parametrize("case", [c for c in W_CASES if f"{c}/dW" in KAT.files])
def test_weight_quantiser_grads(case):
    g = lambda k: KAT[f"{case}/{k}"]
    nb, sym, grp, lwc = [int(v) for v in g("meta")]
    W = t(g("W")).requires_grad_(True)
    up = t(g("up")).requires_grad_(True) if lwc else None
    low = t(g("low")).requires_grad_(True) if lwc else None
    y = orc.fake_quant(W, nb, bool(sym), grp or None, up, low)
    y.backward(t(g("gy")))
    np.testing.assert_allclose(W.grad.numpy(), g("dW"), rtol=1e-6, atol=1e-7, equal_nan=True)
    if lwc:
        np.testing.assert_allclose(up.grad.numpy(), g("d_up"), rtol=1e-5, atol=1e-7, equal_nan=True)
        np.testing.assert_allclose(low.grad.numpy(), g("d_low"), rtol=1e-5, atol=1e-7, equal_nan=True)


@pytest.mark.parametrize("case", A_CASES)
def test_act_quantiser(case):
    g = lambda k: KAT[f"{case}/{k}"]
    nb = int(g("meta")[0])
    X = t(g("X")).requires_grad_(True)
    deq, s, z, _ = orc.fake_quant(X, nb, return_all=True)
    beq(deq.detach().numpy(), g("deq"))
    beq(s.detach().numpy(), g("scale"))
    beq(z.detach().numpy(), g("zp"))
    deq.backward(t(g("gy")))
    np.testing.assert_allclose(X.grad.numpy(), g("dX"), rtol=1e-6, atol=1e-7)
    d2, s2, z2, _ = orc.fake_quant_numpy(g("X").reshape(-1, g("X").shape[-1]), nb, False, None, None, None)
    beq(d2.reshape(g("deq").shape), g("deq"))


def test_survey_kat():
    # SURVEY.md §4 known-answer vector
    assert KAT["kat_pc/zp"].ravel().tolist() == [4.0, 8.0]
    np.testing.assert_allclose(KAT["kat_pc/scale"].ravel(), [0.2618703246116638, 0.09820137917995453], rtol=0, atol=0)
    assert KAT["kat_g4/zp"].ravel().tolist() == [12.0, -3.0, 9.0, 8.0]


def _oracle_setup(fx):
    cfg = orc.QCfg(fx.wbits, fx.abits, fx.group_size, fx.symmetric, fx.lwc, fx.let)
    Wt = fx.weights()
    lin = {"llama": orc.LLAMA_LINEARS, "opt": orc.OPT_LINEARS, "falcon": orc.FALCON_LINEARS}[fx.family]
    P = orc.init_lwc({n: Wt[n] for n in lin}, cfg)
    if fx.let:
        sc, sh = fx.act_scales(), fx.act_shifts()
        oname = "self_attn.o_proj" if fx.family == "llama" else "self_attn.out_proj"
        fname = "mlp.up_proj" if fx.family == "llama" else "fc1"
        for key, mod, w in (("qkv", "self_attn.q_proj", "q"), ("out", oname, "o"), ("fc1", fname, "up" if fx.family == "llama" else "fc1")):
            P[f"{key}_s"] = orc.init_let_scale(sc[mod], Wt[w], 0.5)
            P[f"{key}_d"] = torch.zeros_like(P[f"{key}_s"]) if fx.family == "llama" else sh[mod].clone()
        P["qkt_s"] = torch.ones(Wt["q"].shape[0])
    return cfg, Wt, P


FIXTURES = ["llama_w4a4_lwc_let.npz", "llama_w4a4_lwc_let_bs2.npz", "llama_w3a16g64_lwc.npz",
            "llama_w6a6_lwc_let_aug.npz", "llama_w4a16g64_lwc_let.npz", "opt_w4a16g64_lwc.npz", "opt_w4a4_lwc_let.npz",
            "opt_w4a16_lwc_let.npz", "falcon_w3a16g64_lwc.npz",
            "llama_gqa_w2a16g64_lwc.npz",      # grouped-query attention (Llama-2-70B layout), BASELINE config 4's W2A16g64 LWC
            "llama_sym_w4a16g64_lwc.npz",      # symmetric weight quantiser (main.py --symmetric)
            "opt_w6a6_lwc_let.npz"]


@pytest.mark.parametrize("fname", FIXTURES)
def test_block_trajectory(fname):
    torch.set_num_threads(4)
    fx = BlockFixture(fname)
    cfg, Wt, P = _oracle_setup(fx)
    ref_init = fx.init_params()
    assert set(ref_init) == set(P), (sorted(ref_init), sorted(P))
    for k in P:
        beq(P[k].numpy().ravel(), ref_init[k].numpy().ravel())
    res = orc.calibrate_block(fx.family, Wt, P, cfg, fx.inps, fx.heads, epochs=fx.epochs, batch_size=fx.batch_size,
                              aug_loss=fx.aug_loss, record_grads=True, n_kv_heads=fx.kv_heads)
    np.testing.assert_allclose(res["fp_out"].numpy(), fx.fp_out.numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(np.array(res["losses"]), fx.losses, rtol=2e-3)   # (falcon: SDPA vs explicit softmax re-draws a few 3-bit codes)
    g0 = fx.grads0()
    for k, v in g0.items():
        np.testing.assert_allclose(res["grads"][k].numpy().ravel(), v.numpy().ravel(), rtol=2e-3, atol=1e-9, err_msg=k)
    fin = fx.final_params()
    for k, v in fin.items():
        a, b = res["P"][k].detach().numpy().ravel(), v.numpy().ravel()
        off = np.abs(a - b) > (1e-5 + 1e-3 * np.abs(b))
        # AdamW turns a round-off-sized gradient into a +-lr step: allow a handful of such elements
        assert off.mean() < 5e-3 and np.abs(a - b).max() < 2.5e-2, (k, off.mean(), np.abs(a - b).max())
    Q = fx.quantized()
    lin = {"llama": orc.LLAMA_LINEARS, "opt": orc.OPT_LINEARS, "falcon": orc.FALCON_LINEARS}[fx.family]
    for n in lin:
        w_or, w_ref = res["T_final"][f"{n}.w"].numpy(), Q[f"{n}.w"].numpy()
        frac = np.mean(np.abs(w_or - w_ref) > 1e-6)
        assert frac < 5e-3, (n, frac)      # a handful of codes may flip from 1e-7 differences in trained factors
    np.testing.assert_allclose(res["quant_out"].numpy(), fx.quant_out.numpy(), rtol=0, atol=5e-2)


def test_act_stats_oracle_matches_reference():
    """get_act_scales / get_act_shifts of the reference (generate_act_scale_shift.py) on a toy model: the oracle's
    per-sample statistics folded in sample order reproduce them bit for bit."""
    from golden_util import load
    z = load("act_stats.npz")
    for site, key in (("0", "in/x0"), ("2", "in/x2")):
        xs = torch.from_numpy(z[key])
        per = [orc.act_stats_sample(xs[i]) for i in range(xs.shape[0])]
        scale, shift = orc.act_stats_fold([p[0] for p in per], [p[1] for p in per])
        assert np.array_equal(scale.numpy(), z[f"out/scale/{site}"])
        assert np.array_equal(shift.numpy(), z[f"out/shift/{site}"])


@pytest.mark.parametrize("bits,group", [(4, 128), (3, 128), (2, 64), (8, 0), (3, 0)])
def test_gptq_pack_oracle_round_trip(bits, group):
    """The packer restatement (unpinned third-party format, see oracle/gptq_pack.py): unpack(pack(x)) returns the integer
    codes / zero points of the reference quantiser, and the 3-bit words follow AutoGPTQ's 10+1|10+1|10 split."""
    from oracle import gptq_pack as gp
    rng = np.random.default_rng(bits)
    N, K = 64, 256
    W = torch.from_numpy(rng.standard_normal((N, K)).astype(np.float32) * 0.02)
    up = torch.full((orc.lwc_shape(W, group or None), 1), 4.0)
    deq, s, z, codes = orc.fake_quant(W, bits, False, group or None, up, up.clone(), return_all=True)
    p = gp.pack(deq.numpy(), s.numpy().reshape(N, -1), z.numpy().reshape(N, -1), bits, group)
    c2, z2 = gp.unpack(p, bits, N, K)
    assert np.array_equal(c2, codes.numpy().reshape(N, K).astype(np.int64))
    assert np.array_equal(z2, z.numpy().reshape(N, -1).astype(np.int64))
    assert p["qweight"].shape == (K // 32 * bits, N) and p["qzeros"].shape == (s.numel() // N, N // 32 * bits)
    if bits == 3:      # literal transcription of the published loop for the first 32 input features of column 0
        iw = codes.numpy().reshape(N, K).astype(np.uint32)[0, :32]
        w0 = sum(int(iw[j]) << (3 * j) for j in range(10)) | (int(iw[10]) << 30)
        w1 = ((int(iw[10]) >> 2) & 1) | sum(int(iw[11 + j]) << (3 * j + 1) for j in range(10)) | (int(iw[21]) << 31)
        w2 = ((int(iw[21]) >> 1) & 3) | sum(int(iw[22 + j]) << (3 * j + 2) for j in range(10))
        got = p["qweight"][:3, 0].astype(np.uint32)
        assert [int(x) for x in got] == [w0 & 0xFFFFFFFF, w1 & 0xFFFFFFFF, w2 & 0xFFFFFFFF]
