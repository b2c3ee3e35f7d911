"""Parity of every CUDA kernel (called through the C ABI) against the CPU oracle
(oracle/oq_oracle.py), the golden fixtures generated from the live reference,
and -- for the floating-point GEMM / norm / softmax -- a torch fp32 reference.

Bit-exactness bar (BASELINE.json north_star): integer codes, scales and zero
points of the weight quantiser are compared bit for bit.
"""
import math

import numpy as np
import pytest
import torch

from golden_util import load, t
from oracle import oq_oracle as orc

pytestmark = pytest.mark.gpu

KAT = load("quantizer_kat.npz")
W_CASES = sorted({k.split("/")[0] for k in KAT.files if k.startswith("w")})
A_CASES = sorted({k.split("/")[0] for k in KAT.files if k.startswith("a")})


@pytest.fixture(scope="module")
def ops():
    from omniquant_b200 import ops as o
    return o


def cu(x, dtype=None):
    x = torch.as_tensor(np.asarray(x)) if not isinstance(x, torch.Tensor) else x
    x = x.cuda()
    return x.to(dtype) if dtype is not None else x


def bits_equal(a, b):
    a = np.ascontiguousarray(np.asarray(a, np.float32)).ravel()
    b = np.ascontiguousarray(np.asarray(b, np.float32)).ravel()
    assert a.shape == b.shape
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b)) | ((a == 0) & (b == 0))


def run_wq(ops, W, nb, sym, grp, up, low, recip, **kw):
    N, K = W.shape
    G = math.ceil(K / grp) if grp else 1
    out = dict(wq_f32=torch.empty(N, K, device="cuda"), wq_bf16=torch.empty(N, K, device="cuda", dtype=torch.bfloat16),
               codes=torch.empty(N, K, device="cuda", dtype=torch.uint8),
               scale_out=torch.empty(N * G, device="cuda"), zp_out=torch.empty(N * G, device="cuda"),
               xmin_out=torch.empty(N * G, device="cuda"), xmax_out=torch.empty(N * G, device="cuda"))
    ops.weight_quant_fwd(W, n_bits=nb, symmetric=sym, group_size=grp, up=up, low=low, recip_div=recip, **out, **kw)
    torch.cuda.synchronize()
    return out


# ---------------------------------------------------------------- weight quantiser: bit-exact
@pytest.mark.parametrize("case", W_CASES)
def test_weight_quant_bit_exact_vs_golden(ops, case):
    """recip_div=0 (CPU ATen division semantics) against fixtures made by the reference on CPU.
    The kernel evaluates sigmoid() itself (CUDA expf) while the fixture used ATen's CPU sigmoid, so
    first check with the numpy oracle fed the *device* sigmoid (must be bit-exact everywhere), then
    against the fixture directly (bit-exact wherever the two sigmoids agree bit for bit)."""
    g = lambda k: KAT[f"{case}/{k}"]
    nb, sym, grp, lwc = [int(v) for v in g("meta")]
    W = cu(g("W"))
    up = cu(g("up")).reshape(-1) if lwc else None
    low = cu(g("low")).reshape(-1) if lwc else None
    out = run_wq(ops, W, nb, bool(sym), grp, up, low, False)
    sig_up = torch.sigmoid(up).cpu().numpy() if lwc else None
    sig_low = torch.sigmoid(low).cpu().numpy() if lwc else None
    deq, s, z, codes = orc.fake_quant_numpy(g("W"), nb, bool(sym), grp or None, sig_up, sig_low)
    N, K = g("W").shape
    assert bits_equal(out["scale_out"].cpu().numpy(), s).all()
    assert bits_equal(out["zp_out"].cpu().numpy(), z).all()
    assert bits_equal(out["wq_f32"].cpu().numpy(), deq).all()
    cg = codes.reshape(N, -1)[:, :K]
    ok = np.isnan(cg) | (cg == out["codes"].cpu().numpy().astype(np.float32))
    assert ok.all()
    # bf16 operand = round-to-nearest-even of the fp32 dequantised value
    ref_bf = torch.from_numpy(deq).to(torch.bfloat16)
    got_bf = out["wq_bf16"].cpu()
    assert (torch.isnan(ref_bf) | (ref_bf == got_bf)).all()
    # straight against the reference's fixture
    same_sig = np.ones(s.size, bool)
    if lwc and f"{case}/sig_up" in KAT.files:
        same_sig = bits_equal(sig_up, g("sig_up")) & bits_equal(sig_low, g("sig_low"))
    se = bits_equal(out["scale_out"].cpu().numpy(), g("scale"))
    assert se[same_sig].all()
    assert bits_equal(out["zp_out"].cpu().numpy(), g("zp"))[same_sig].all()
    frac = 1.0 - bits_equal(out["wq_f32"].cpu().numpy(), g("deq")).mean()
    assert frac <= (1.0 - same_sig.mean()) + 1e-12, f"dequant mismatch fraction {frac}"


@pytest.mark.parametrize("case", W_CASES)
def test_weight_quant_bit_exact_vs_cuda_aten(ops, case):
    """recip_div=1: the reference's own arithmetic executed by ATen on this GPU (the oracle's torch
    functions run on CUDA tensors) -- scales, zero points, codes bit-exact."""
    g = lambda k: KAT[f"{case}/{k}"]
    nb, sym, grp, lwc = [int(v) for v in g("meta")]
    W = cu(g("W"))
    up = cu(g("up")) if lwc else None
    low = cu(g("low")) if lwc else None
    deq, s, z, codes = orc.fake_quant(W, nb, bool(sym), grp or None, up, low, return_all=True)
    out = run_wq(ops, W, nb, bool(sym), grp, up.reshape(-1) if lwc else None, low.reshape(-1) if lwc else None, True)
    assert bits_equal(out["scale_out"].cpu().numpy(), s.cpu().numpy()).all()
    assert bits_equal(out["zp_out"].cpu().numpy(), z.cpu().numpy()).all()
    assert bits_equal(out["wq_f32"].cpu().numpy(), deq.cpu().numpy()).all()


def test_weight_quant_large_bit_exact(ops):
    """LLaMA-7B sized rows (K=4096 and 11008, per-channel and g128), random LWC factors."""
    gen = torch.Generator(device="cuda").manual_seed(7)
    for (N, K, grp, nb) in [(512, 4096, 0, 4), (256, 11008, 0, 4), (256, 4096, 128, 3), (128, 11008, 64, 2)]:
        W = (torch.randn(N, K, device="cuda", generator=gen) * 0.02).half().float()
        G = K // grp if grp else 1
        up = torch.randn(N * G, 1, device="cuda", generator=gen) * 1.5 + 3
        low = torch.randn(N * G, 1, device="cuda", generator=gen) * 1.5 + 3
        deq, s, z, _ = orc.fake_quant(W, nb, False, grp or None, up, low, return_all=True)
        out = run_wq(ops, W, nb, False, grp, up.reshape(-1), low.reshape(-1), True)
        assert torch.equal(out["scale_out"], s.reshape(-1))
        assert torch.equal(out["zp_out"], z.reshape(-1))
        assert torch.equal(out["wq_f32"], deq)


def test_weight_quant_full_size_properties(ops):
    """Size-independent properties at LLaMA-7B weight sizes (no oracle needed): power-of-two scaling leaves codes
    and zero points untouched and scales `scale` / dequant exactly; codes in range; dequant == (code - zp) * scale."""
    gen = torch.Generator(device="cuda").manual_seed(21)
    for (N, K, grp, nb) in [(4096, 11008, 0, 4), (11008, 4096, 128, 3), (4096, 4096, 64, 2)]:
        W = (torch.randn(N, K, device="cuda", generator=gen) * 0.02).half().float()
        G = K // grp if grp else 1
        up = torch.randn(N * G, device="cuda", generator=gen) + 3
        low = torch.randn(N * G, device="cuda", generator=gen) + 3
        a = run_wq(ops, W, nb, False, grp, up, low, True)
        b = run_wq(ops, W * 4.0, nb, False, grp, up, low, True)
        assert torch.equal(a["codes"], b["codes"])
        assert torch.equal(a["zp_out"], b["zp_out"])
        assert torch.equal(a["scale_out"] * 4.0, b["scale_out"])
        assert torch.equal(a["wq_f32"] * 4.0, b["wq_f32"])
        assert int(a["codes"].max()) <= 2 ** nb - 1
        rep = lambda v: v.view(N, G, 1).expand(N, G, K // G).reshape(N, K)
        assert torch.equal((a["codes"].float() - rep(a["zp_out"])) * rep(a["scale_out"]), a["wq_f32"])
        # every level of the 2^n grid is used somewhere
        assert len(torch.unique(a["codes"])) == 2 ** nb


def test_weight_quant_let_fold_and_matvec(ops):
    """fold (W*col)/row and (W*col)*row in the reference's op order + W@shift."""
    gen = torch.Generator(device="cuda").manual_seed(3)
    N, K = 192, 320
    W = (torch.randn(N, K, device="cuda", generator=gen) * 0.05).half().float()
    col = torch.rand(K, device="cuda", generator=gen) + 0.5
    row = torch.rand(N, device="cuda", generator=gen) + 0.5
    shift = torch.randn(K, device="cuda", generator=gen) * 0.1
    up = torch.full((N, 1), 4.0, device="cuda")
    for grp in (0, 64):
        G = K // grp if grp else 1
        upg = torch.full((N * G, 1), 4.0, device="cuda")
        for row_op, fold in ((2, lambda w: (w * col.view(1, -1)) / row.view(-1, 1)),
                             (1, lambda w: (w * col.view(1, -1)) * row.view(-1, 1)), (0, lambda w: w * col.view(1, -1))):
            mv = torch.empty(N, device="cuda")
            out = run_wq(ops, W, 4, False, grp, upg.reshape(-1), upg.reshape(-1), True, col_scale=col,
                         row_scale=row if row_op else None, row_op=row_op, shift=shift, matvec_out=mv)
            deq, s, z, _ = orc.fake_quant(fold(W), 4, False, grp or None, upg, upg, return_all=True)
            assert torch.equal(out["wq_f32"], deq)
            torch.testing.assert_close(mv, W @ shift, rtol=1e-4, atol=1e-5)


# ---------------------------------------------------------------- weight quantiser backward
@pytest.mark.parametrize("case", [c for c in W_CASES if f"{c}/dW" in KAT.files and c != "w4_pc_asym"])
def test_weight_quant_bwd_vs_golden(ops, case):
    g = lambda k: KAT[f"{case}/{k}"]
    nb, sym, grp, lwc = [int(v) for v in g("meta")]
    W = cu(g("W"))
    N, K = W.shape
    up = cu(g("up")).reshape(-1) if lwc else None
    low = cu(g("low")).reshape(-1) if lwc else None
    out = run_wq(ops, W, nb, bool(sym), grp, up, low, False)
    G = out["scale_out"].numel()
    d_up, d_low = torch.zeros(G, device="cuda"), torch.zeros(G, device="cuda")
    dW = torch.zeros(N, K, device="cuda")
    ops.weight_quant_bwd(W, cu(g("gy")), out["xmin_out"], out["xmax_out"], n_bits=nb, symmetric=bool(sym),
                         group_size=grp, up=up, low=low, recip_div=False, d_up=d_up, d_low=d_low, dW_full=dW)
    torch.cuda.synchronize()
    np.testing.assert_allclose(dW.cpu().numpy(), g("dW"), rtol=1e-4, atol=1e-5)
    if lwc:
        np.testing.assert_allclose(d_up.cpu().numpy(), g("d_up").ravel(), rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(d_low.cpu().numpy(), g("d_low").ravel(), rtol=2e-4, atol=2e-5)


@pytest.mark.parametrize("grp,row_op", [(0, 2), (0, 1), (64, 2), (0, 0)])
def test_weight_quant_bwd_let_reductions(ops, grp, row_op):
    """dcol / drow / dshift / d_up / d_low against oracle autograd through fold + fake-quant."""
    gen = torch.Generator().manual_seed(11)
    N, K = 96, 256
    W = (torch.randn(N, K, generator=gen) * 0.05).half().float()
    col = (torch.rand(K, generator=gen) + 0.5).requires_grad_(True)
    row = (torch.rand(N, generator=gen) + 0.5).requires_grad_(True)
    shift = (torch.randn(K, generator=gen) * 0.1).requires_grad_(True)
    G = K // grp if grp else 1
    up = (torch.randn(N * G, 1, generator=gen) + 3).requires_grad_(True)
    low = (torch.randn(N * G, 1, generator=gen) + 3).requires_grad_(True)
    gy = torch.randn(N, K, generator=gen)
    e = torch.randn(N, generator=gen)
    W1 = W * col.view(1, -1)
    W2 = W1 / row.view(-1, 1) if row_op == 2 else (W1 * row.view(-1, 1) if row_op == 1 else W1)
    y = orc.fake_quant(W2, 4, False, grp or None, up, low)
    ((y * gy).sum() + ((W @ shift) * e).sum()).backward()
    Wc = W.cuda()
    out = run_wq(ops, Wc, 4, False, grp, up.detach().cuda().reshape(-1), low.detach().cuda().reshape(-1), False,
                 col_scale=col.detach().cuda(), row_scale=row.detach().cuda() if row_op else None, row_op=row_op)
    d_up, d_low = torch.zeros(N * G, device="cuda"), torch.zeros(N * G, device="cuda")
    dcol, dshift, drow = torch.zeros(K, device="cuda"), torch.zeros(K, device="cuda"), torch.zeros(N, device="cuda")
    for gdt in (torch.float32, torch.bfloat16):
        dcol.zero_(); dshift.zero_()
        ops.weight_quant_bwd(Wc, gy.cuda().to(gdt), out["xmin_out"], out["xmax_out"], n_bits=4, group_size=grp,
                             up=up.detach().cuda().reshape(-1), low=low.detach().cuda().reshape(-1), recip_div=False,
                             col_scale=col.detach().cuda(), row_scale=row.detach().cuda() if row_op else None,
                             row_op=row_op, e_row=e.cuda(), d_up=d_up, d_low=d_low, dcol_acc=dcol, drow=drow,
                             dshift_acc=dshift)
        torch.cuda.synchronize()
        tol = dict(rtol=2e-3, atol=2e-3) if gdt == torch.float32 else dict(rtol=3e-2, atol=5e-2)
        torch.testing.assert_close(dcol.cpu(), col.grad, **tol)
        torch.testing.assert_close(dshift.cpu(), shift.grad, rtol=1e-3, atol=1e-4)
        torch.testing.assert_close(d_up.cpu(), up.grad.reshape(-1), **tol)
        torch.testing.assert_close(d_low.cpu(), low.grad.reshape(-1), **tol)
        if row_op:
            # drow holds sum_k dW2*W1; apply the op's own factor
            r = row.detach()
            ref = drow.cpu() * (1.0 if row_op == 1 else -1.0) / (1.0 if row_op == 1 else r * r)
            torch.testing.assert_close(ref, row.grad, **tol)


# ---------------------------------------------------------------- activation quantiser
@pytest.mark.parametrize("case", A_CASES)
def test_act_quant_vs_golden(ops, case):
    """bf16-representable inputs: fixtures hold fp16 values, so re-run the oracle on the bf16-rounded
    input (fp32 arithmetic, CPU division semantics) and compare dequantised values exactly."""
    g = lambda k: KAT[f"{case}/{k}"]
    nb = int(g("meta")[0])
    X = t(g("X")).to(torch.bfloat16)
    C = X.shape[-1]
    Xf = X.float().reshape(-1, C).requires_grad_(True)
    deq, s, z, _ = orc.fake_quant(Xf, nb, return_all=True)
    gy = t(g("gy")).to(torch.bfloat16).reshape(-1, C)
    deq.backward(gy.float())
    x = X.reshape(-1, C).cuda()
    y = torch.empty_like(x)
    sc, zp = torch.empty(x.shape[0], device="cuda"), torch.empty(x.shape[0], device="cuda")
    ops.act_quant_fwd(x, y, n_bits=nb, recip_div=False, scale_out=sc, zp_out=zp)
    gx = torch.empty_like(x)
    ops.act_quant_bwd(x, gy.cuda(), gx, n_bits=nb, recip_div=False)
    torch.cuda.synchronize()
    assert torch.equal(sc.cpu(), s.detach().reshape(-1))
    assert torch.equal(zp.cpu(), z.detach().reshape(-1))
    assert torch.equal(y.cpu(), deq.detach().to(torch.bfloat16))
    torch.testing.assert_close(gx.cpu().float(), Xf.grad, rtol=1e-2, atol=1e-2)


def test_act_quant_rope_heads(ops):
    """fused rotary + per-(token, head) quant on the strided [T, 3H] QKV buffer (d=128 and d=64)."""
    gen = torch.Generator().manual_seed(5)
    for d, h, T in ((128, 4, 48), (64, 6, 40)):
        H = h * d
        qkv = torch.randn(T, 3 * H, generator=gen).to(torch.bfloat16)
        cos, sin = orc.rope_tables(d, T)
        q = qkv[:, :H].float().view(T, h, d).requires_grad_(True)
        qr, _ = orc.apply_rope(q.transpose(0, 1)[None], q.transpose(0, 1)[None], cos, sin)
        ref = orc.fake_quant(qr, 4)                      # [1,h,T,d]
        gy = torch.randn(T, h, d, generator=gen).to(torch.bfloat16)
        ref.backward(gy.float().transpose(0, 1)[None])
        dq = qkv.cuda()
        x = dq[:, :H].view(T, h, d)
        y = torch.empty(T, h, d, device="cuda", dtype=torch.bfloat16)
        ops.act_quant_fwd(x, y, n_bits=4, recip_div=False, cos=cos.cuda(), sin=sin.cuda(), seq_len=T)
        gx = torch.empty(T, h, d, device="cuda", dtype=torch.bfloat16)
        ops.act_quant_bwd(x, gy.cuda(), gx, n_bits=4, recip_div=False, cos=cos.cuda(), sin=sin.cuda(), seq_len=T)
        torch.cuda.synchronize()
        refy = ref.detach()[0].transpose(0, 1)
        # codes can flip where the fused fp32 rope differs in the last ulp from the unfused one
        mism = (y.cpu().float() - refy).abs() > 1e-2 * refy.abs().max()
        assert mism.float().mean() < 2e-3
        torch.testing.assert_close(gx.cpu().float(), q.grad, rtol=5e-2, atol=5e-2)
        # rope only
        y2 = torch.empty_like(y)
        ops.act_quant_fwd(x, y2, n_bits=16, cos=cos.cuda(), sin=sin.cuda(), seq_len=T)
        torch.testing.assert_close(y2.cpu().float(), qr.detach()[0].transpose(0, 1), rtol=1e-2, atol=1e-2)


def test_act_quant_fused_qkv_launch(ops):
    """One launch over [q heads | k heads | v heads] of a token (rope_inner = rotated heads) must equal the three
    separate launches bit for bit, forward and backward, for MHA and GQA head counts (chunk tails included) and for
    the rotary-only (n_bits=16) mode, where the V heads are copied."""
    gen = torch.Generator().manual_seed(11)
    for d, h, hkv, T in ((128, 5, 5, 37), (64, 12, 3, 29), (128, 32, 32, 16)):
        nh = h + 2 * hkv
        qkv = torch.randn(T, nh * d, generator=gen).to(torch.bfloat16).cuda()
        gy = torch.randn(T, nh * d, generator=gen).to(torch.bfloat16).cuda()
        cos, sin = (v.cuda() for v in orc.rope_tables(d, T))
        for nb in (4, 16):
            y1, y3 = torch.zeros_like(qkv), torch.zeros_like(qkv)
            g1, g3 = torch.zeros_like(qkv), torch.zeros_like(qkv)
            sc1, sc3 = torch.zeros(T * nh, device="cuda"), torch.zeros(T, nh, device="cuda")
            v3 = lambda t_: t_.view(T, nh, d)
            kw = dict(scale_out=sc1) if nb < 16 else {}
            ops.act_quant_fwd(v3(qkv), v3(y1), n_bits=nb, recip_div=True, cos=cos, sin=sin, seq_len=T, rope_inner=h + hkv, **kw)
            ops.act_quant_bwd(v3(qkv), v3(gy), v3(g1), n_bits=nb, recip_div=True, cos=cos, sin=sin, seq_len=T,
                              rope_inner=h + hkv)
            for a, b, rope in ((0, h + hkv, True), (h + hkv, nh, False)):
                if not rope and nb >= 16:
                    v3(y3)[:, a:b].copy_(v3(qkv)[:, a:b])
                    v3(g3)[:, a:b].copy_(v3(gy)[:, a:b])
                    continue
                rk = dict(cos=cos, sin=sin, seq_len=T) if rope else {}
                sc_part = torch.zeros(T * (b - a), device="cuda")
                fk = dict(scale_out=sc_part) if nb < 16 else {}
                ops.act_quant_fwd(v3(qkv)[:, a:b], v3(y3)[:, a:b], n_bits=nb, recip_div=True, **rk, **fk)
                ops.act_quant_bwd(v3(qkv)[:, a:b], v3(gy)[:, a:b], v3(g3)[:, a:b], n_bits=nb, recip_div=True, **rk)
                sc3[:, a:b] = sc_part.view(T, b - a)
            torch.cuda.synchronize()
            assert torch.equal(y1, y3) and torch.equal(g1, g3)
            if nb < 16:
                assert torch.equal(sc1.view(T, nh), sc3)


@pytest.mark.parametrize("F", [11008, 3072, 256, 16384])
def test_swiglu_act_quant_fused_equals_unfused(ops, F):
    """oq_swiglu_act_quant_fwd/bwd == oq_swiglu_fwd -> oq_act_quant_fwd (and the two backwards) bit for bit."""
    gen = torch.Generator().manual_seed(F)
    rows = 19
    gu = (torch.randn(rows, 2 * F, generator=gen) * 1.5).to(torch.bfloat16).cuda()
    gy = torch.randn(rows, F, generator=gen).to(torch.bfloat16).cuda()
    for nb in (4, 6):
        m, mq, mq2 = (torch.empty(rows, F, device="cuda", dtype=torch.bfloat16) for _ in range(3))
        ops.swiglu_fwd(gu, m)
        ops.act_quant_fwd(m, mq, n_bits=nb, recip_div=True)
        ops.swiglu_act_quant_fwd(gu, mq2, n_bits=nb, recip_div=True)
        dm = torch.empty_like(m)
        d1, d2 = torch.empty_like(gu), torch.empty_like(gu)
        ops.act_quant_bwd(m, gy, dm, n_bits=nb, recip_div=True)
        ops.swiglu_bwd(gu, dm, d1)
        ops.swiglu_act_quant_bwd(gu, gy, d2, n_bits=nb, recip_div=True)
        torch.cuda.synchronize()
        assert torch.equal(mq, mq2)
        assert torch.equal(d1, d2)


@pytest.mark.parametrize("C,kind", [(4096, 0), (768, 1), (8192, 0), (14848, 1), (256, 0)])
def test_norm_act_quant_fused_equals_unfused(ops, C, kind):
    """oq_norm_act_quant_fwd == oq_norm_fwd -> oq_act_quant_fwd bit for bit (y, yq, row statistics)."""
    gen = torch.Generator().manual_seed(C + kind)
    rows = 21
    x = (torch.randn(rows, C, generator=gen) * 2).to(torch.bfloat16).cuda()
    w = (1 + 0.1 * torch.randn(C, generator=gen)).cuda()
    b = (0.1 * torch.randn(C, generator=gen)).cuda() if kind == 1 else None
    y1, y2, q1, q2 = (torch.empty_like(x) for _ in range(4))
    st = [torch.zeros(rows, device="cuda") for _ in range(4)]
    ops.norm_fwd(x, w, b, 1e-5, kind, y1, st[0], st[1])
    ops.act_quant_fwd(y1, q1, n_bits=4, recip_div=True)
    ops.norm_act_quant_fwd(x, w, b, 1e-5, kind, y2, q2, n_bits=4, recip_div=True, mean_out=st[2], rstd_out=st[3])
    q3 = torch.empty_like(x)
    ops.norm_act_quant_fwd(x, w, b, 1e-5, kind, None, q3, n_bits=4, recip_div=True)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2) and torch.equal(q1, q2) and torch.equal(q1, q3)
    assert torch.equal(st[0], st[2]) and torch.equal(st[1], st[3])


def test_act_quant_large_rows(ops):
    gen = torch.Generator().manual_seed(9)
    for C in (4096, 11008, 768, 100):
        X = (torch.randn(64, C, generator=gen) * 2).to(torch.bfloat16)
        X[:, 7] *= 20
        ref, s, z, _ = orc.fake_quant(X.float(), 4, return_all=True)
        x = X.cuda()
        y = torch.empty_like(x)
        sc = torch.empty(64, device="cuda")
        ops.act_quant_fwd(x, y, n_bits=4, recip_div=False, scale_out=sc)
        torch.cuda.synchronize()
        assert torch.equal(sc.cpu(), s.reshape(-1))
        assert torch.equal(y.cpu(), ref.to(torch.bfloat16))


# ---------------------------------------------------------------- norms
@pytest.mark.parametrize("kind,C", [(0, 4096), (0, 256), (1, 768), (1, 100)])
def test_norm_fwd_bwd(ops, kind, C):
    gen = torch.Generator().manual_seed(2)
    R = 96
    X = torch.randn(R, C, generator=gen).to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn(C, generator=gen)).requires_grad_(True)
    b = (0.1 * torch.randn(C, generator=gen)).requires_grad_(True)
    xf = X.float().requires_grad_(True)
    eps = 1e-6 if kind == 0 else 1e-5
    ref = orc.rms_norm(xf, w, b, eps) if kind == 0 else orc.layer_norm(xf, w, b, eps)
    gy = torch.randn(R, C, generator=gen).to(torch.bfloat16)
    add = torch.randn(R, C, generator=gen).to(torch.bfloat16)
    ref.backward(gy.float())
    x = X.cuda()
    y = torch.empty_like(x)
    mean, rstd = torch.empty(R, device="cuda"), torch.empty(R, device="cuda")
    ops.norm_fwd(x, w.detach().cuda(), b.detach().cuda(), eps, kind, y, mean, rstd)
    dx = torch.empty_like(x)
    dw, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.norm_bwd(x, w.detach().cuda(), kind, mean, rstd, gy.cuda(), dx=dx, dx_add=add.cuda(), dw_acc=dw, db_acc=db)
    torch.cuda.synchronize()
    torch.testing.assert_close(y.cpu().float(), ref.detach(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(dx.cpu().float(), xf.grad + add.float(), rtol=2e-2, atol=2e-2)
    torch.testing.assert_close(dw.cpu(), w.grad, rtol=1e-3, atol=1e-2)
    torch.testing.assert_close(db.cpu(), b.grad, rtol=1e-3, atol=1e-2)


# ---------------------------------------------------------------- GEMM (tcgen05)
def _gemm_ref(A, B, alpha=1.0, bias=None, res=None):
    out = alpha * (A.float() @ B.float().transpose(-1, -2))
    if bias is not None:
        out = out + bias
    if res is not None:
        out = out + res.float()
    return out


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K,tile_n", [(256, 512, 256, 256), (200, 344, 136, 0), (128, 128, 64, 128), (384, 264, 1032, 256)])
def test_gemm_layouts(ops, a_mn, b_mn, M, N, K, tile_n):
    gen = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=gen).to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda", generator=gen).to(torch.bfloat16)
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    bias = torch.randn(N, device="cuda", generator=gen)
    res = torch.randn(M, N, device="cuda", generator=gen).to(torch.bfloat16)
    for cdt in (torch.float32, torch.bfloat16, torch.float16):
        C = torch.full((M, N), float("nan"), device="cuda", dtype=cdt)
        ops.gemm(As, Bs, C, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, lda=As.stride(0), ldb=Bs.stride(0), ldc=N, bias=bias,
                 residual=res, ldr=N, alpha=0.5, tile_n=tile_n)
        torch.cuda.synchronize()
        ref = _gemm_ref(A, B, 0.5, bias, res)
        tol = dict(rtol=1e-3, atol=1e-2) if cdt == torch.float32 else (dict(rtol=2e-2, atol=1e-1) if cdt == torch.bfloat16
                                                                       else dict(rtol=3e-3, atol=2e-2))
        torch.testing.assert_close(C.float(), ref, **tol)


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", [(2000, 5000, 1000), (2048, 2560, 192), (1160, 4104, 4096)])
def test_gemm_cta_pair_kernel(ops, a_mn, b_mn, M, N, K):
    """Shapes with at least one 256 x 256 tile per SM pair go through the cta_group::2 kernel (two CTAs per tile, each
    staging its own A rows and half of B): ragged M / N / K against the fp32 product, every operand layout and output type."""
    gen = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=gen).to(torch.bfloat16)
    B = torch.randn(N, K, device="cuda", generator=gen).to(torch.bfloat16)
    As = A.t().contiguous() if a_mn else A
    Bs = B.t().contiguous() if b_mn else B
    bias = torch.randn(N, device="cuda", generator=gen)
    res = torch.randn(M, N, device="cuda", generator=gen).to(torch.bfloat16)
    ref = _gemm_ref(A, B, 0.25, bias, res)
    for cdt in (torch.float32, torch.bfloat16):
        C = torch.full((M, N), float("nan"), device="cuda", dtype=cdt)
        ops.gemm(As, Bs, C, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, lda=As.stride(0), ldb=Bs.stride(0), ldc=N, bias=bias,
                 residual=res, ldr=N, alpha=0.25)
        torch.cuda.synchronize()
        tol = dict(rtol=1e-3, atol=2e-2) if cdt == torch.float32 else dict(rtol=2e-2, atol=2e-1)
        torch.testing.assert_close(C.float(), ref, **tol)
    # and the same product through the single-CTA kernel (tile_n = 128 keeps it there): identical fp32 sums
    C1 = torch.empty(M, N, device="cuda")
    C2 = torch.empty(M, N, device="cuda")
    ops.gemm(As, Bs, C1, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, lda=As.stride(0), ldb=Bs.stride(0), ldc=N)
    ops.gemm(As, Bs, C2, M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, lda=As.stride(0), ldb=Bs.stride(0), ldc=N, tile_n=128)
    torch.cuda.synchronize()
    torch.testing.assert_close(C1, C2, rtol=1e-5, atol=1e-3)


def test_gemm_llama_shapes_and_linear_helpers(ops):
    gen = torch.Generator(device="cuda").manual_seed(1)
    M, K, N = 2048, 4096, 4096
    x = torch.randn(M, K, device="cuda", generator=gen).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=gen) * 0.02).to(torch.bfloat16)
    dy = torch.randn(M, N, device="cuda", generator=gen).to(torch.bfloat16)
    y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.linear_fwd(x, w, y)
    dx = torch.empty(M, K, device="cuda", dtype=torch.bfloat16)
    ops.linear_dgrad(dy, w, dx)
    dw = torch.empty(N, K, device="cuda", dtype=torch.float32)
    ops.linear_wgrad(dy, x, dw)
    torch.cuda.synchronize()
    torch.testing.assert_close(y.float(), x.float() @ w.float().t(), rtol=2e-2, atol=5e-2)
    torch.testing.assert_close(dx.float(), dy.float() @ w.float(), rtol=2e-2, atol=5e-2)
    torch.testing.assert_close(dw, dy.float().t() @ x.float(), rtol=1e-3, atol=5e-2)


def test_gemm_batched_heads_and_causal(ops):
    """QuantMatMul shapes: per-head operands are strided slices of [T, H] buffers."""
    gen = torch.Generator(device="cuda").manual_seed(4)
    T, h, d = 384, 3, 128
    H = h * d
    Q = torch.randn(T, H, device="cuda", generator=gen).to(torch.bfloat16)
    Kt = torch.randn(T, H, device="cuda", generator=gen).to(torch.bfloat16)
    V = torch.randn(T, H, device="cuda", generator=gen).to(torch.bfloat16)
    qh = Q.view(T, h, d).transpose(0, 1).float()
    kh = Kt.view(T, h, d).transpose(0, 1).float()
    vh = V.view(T, h, d).transpose(0, 1).float()
    tri = torch.tril(torch.ones(T, T, device="cuda", dtype=torch.bool))
    for causal in (0, 1):
        S = torch.zeros(h, T, T, device="cuda")
        ops.gemm(Q, Kt, S, M=T, N=T, K=d, lda=H, ldb=H, ldc=T, batch=h, strideA=d, strideB=d, strideC=T * T,
                 alpha=1 / math.sqrt(d), causal=causal)
        torch.cuda.synchronize()
        ref = qh @ kh.transpose(1, 2) / math.sqrt(d)
        if causal:
            torch.testing.assert_close(S[:, tri], ref[:, tri], rtol=1e-3, atol=1e-2)
        else:
            torch.testing.assert_close(S, ref, rtol=1e-3, atol=1e-2)
    P = torch.softmax(ref.masked_fill(~tri, float("-inf")), -1).to(torch.bfloat16)
    # O = P V  (B operand N-major), k-limited
    for causal in (0, 2):
        O = torch.empty(T, H, device="cuda", dtype=torch.bfloat16)
        ops.gemm(P, V, O, M=T, N=d, K=T, lda=T, ldb=H, b_mn=True, ldc=H, batch=h, strideA=T * T, strideB=d, strideC=d,
                 causal=causal)
        torch.cuda.synchronize()
        refO = (P.float() @ vh).transpose(0, 1).reshape(T, H)
        torch.testing.assert_close(O.float(), refO, rtol=2e-2, atol=2e-2)
    # dV = P^T dO (both MN-major), k-start limited
    dO = torch.randn(T, H, device="cuda", generator=gen).to(torch.bfloat16)
    for causal in (0, 3):
        dV = torch.empty(T, H, device="cuda", dtype=torch.bfloat16)
        ops.gemm(P, dO, dV, M=T, N=d, K=T, a_mn=True, lda=T, b_mn=True, ldb=H, ldc=H, batch=h, strideA=T * T, strideB=d,
                 strideC=d, causal=causal)
        torch.cuda.synchronize()
        refdV = (P.float().transpose(1, 2) @ dO.view(T, h, d).transpose(0, 1).float()).transpose(0, 1).reshape(T, H)
        torch.testing.assert_close(dV.float(), refdV, rtol=2e-2, atol=5e-2)


# ---------------------------------------------------------------- elementwise
def test_swiglu_relu(ops):
    gen = torch.Generator().manual_seed(6)
    R, Fd = 40, 704
    gu = torch.randn(R, 2 * Fd, generator=gen).to(torch.bfloat16)
    g = gu[:, :Fd].float().requires_grad_(True)
    u = gu[:, Fd:].float().requires_grad_(True)
    ref = torch.nn.functional.silu(g) * u
    dm = torch.randn(R, Fd, generator=gen).to(torch.bfloat16)
    ref.backward(dm.float())
    m = torch.empty(R, Fd, device="cuda", dtype=torch.bfloat16)
    ops.swiglu_fwd(gu.cuda(), m)
    dgu = torch.empty(R, 2 * Fd, device="cuda", dtype=torch.bfloat16)
    ops.swiglu_bwd(gu.cuda(), dm.cuda(), dgu)
    torch.cuda.synchronize()
    torch.testing.assert_close(m.cpu().float(), ref.detach(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(dgu.cpu().float(), torch.cat([g.grad, u.grad], 1), rtol=2e-2, atol=2e-2)
    x = torch.randn(R, Fd, generator=gen).to(torch.bfloat16).cuda()
    y, dx = torch.empty_like(x), torch.empty_like(x)
    ops.relu_fwd(x, y)
    ops.relu_bwd(x, dm.cuda(), dx)
    torch.cuda.synchronize()
    assert torch.equal(y, torch.relu(x))
    assert torch.equal(dx, torch.where(x > 0, dm.cuda(), torch.zeros_like(x)))


def test_softmax_fwd_bwd(ops):
    gen = torch.Generator().manual_seed(8)
    B, T = 3, 200
    _softmax_case(ops, 2, 2048)
    S = (torch.randn(B, T, T, generator=gen) * 3).requires_grad_(True)
    mask = orc.causal_mask(T)[0]
    ref = torch.softmax(torch.max(S + mask, torch.tensor(torch.finfo(torch.float32).min)), -1)
    dP = torch.randn(B, T, T, generator=gen)
    ref.backward(dP)
    for kw in (dict(causal=True), dict(mask=mask.cuda())):
        # causal rows are only written up to the end of their 128-wide diagonal tile (the k-limited GEMMs read no
        # further): start from zeros so the untouched tail compares equal to the reference's exact zeros
        P = torch.zeros(B, T, T, device="cuda", dtype=torch.bfloat16)
        ops.softmax_fwd(S.detach().cuda(), P, **kw)
        dS = torch.zeros_like(P)
        ops.softmax_bwd(P, dP.cuda(), dS, causal=kw.get("causal", False))
        torch.cuda.synchronize()
        torch.testing.assert_close(P.cpu().float(), ref.detach(), rtol=1e-2, atol=4e-3)
        torch.testing.assert_close(dS.cpu().float(), S.grad, rtol=5e-2, atol=1e-2)


def _softmax_case(ops, B, T):
    gen = torch.Generator(device="cuda").manual_seed(T)
    S = torch.randn(B, T, T, device="cuda", generator=gen) * 3
    dP = torch.randn(B, T, T, device="cuda", generator=gen)
    tri = torch.tril(torch.ones(T, T, device="cuda", dtype=torch.bool))
    Sr = S.clone().requires_grad_(True)
    ref = torch.softmax(Sr.masked_fill(~tri, float("-inf")), -1)
    ref.backward(dP)
    for sdt, gdt in ((torch.float32, torch.float32), (torch.float16, torch.bfloat16)):   # engine uses fp16 scores, bf16 dP
        P = torch.zeros(B, T, T, device="cuda", dtype=torch.bfloat16)
        ops.softmax_fwd(S.to(sdt), P, causal=True)
        dS = torch.zeros_like(P)
        ops.softmax_bwd(P, dP.to(gdt), dS, causal=True)
        torch.cuda.synchronize()
        torch.testing.assert_close(P.float(), ref.detach(), rtol=2e-2, atol=4e-3)
        torch.testing.assert_close(dS.float(), Sr.grad, rtol=5e-2, atol=2e-2)


def test_mse_adamw_colsum_truncate_transpose(ops):
    gen = torch.Generator().manual_seed(10)
    n = 64 * 256
    out = torch.randn(64, 256, generator=gen).to(torch.bfloat16)
    tgt = torch.randn(64, 256, generator=gen).to(torch.bfloat16)
    tg2 = torch.randn(64, 256, generator=gen).to(torch.bfloat16)
    o = out.float().requires_grad_(True)
    ref = torch.nn.functional.mse_loss(tgt.float(), o) + torch.nn.functional.mse_loss(tg2.float(), o)
    ref.backward()
    loss = torch.zeros(1, device="cuda")
    dout = torch.empty(64, 256, device="cuda", dtype=torch.bfloat16)
    ops.mse_fwd_bwd(out.cuda(), tgt.cuda(), loss, dout, target2=tg2.cuda())
    torch.cuda.synchronize()
    torch.testing.assert_close(loss.cpu()[0], ref.detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(dout.cpu().float(), o.grad, rtol=1e-2, atol=1e-7)
    # AdamW: 5 steps against torch.optim.AdamW
    p0 = torch.randn(1000, generator=gen)
    pt = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([pt], lr=1e-2, weight_decay=0.1)
    p = p0.cuda()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    for _ in range(5):
        g = torch.randn(1000, generator=gen)
        pt.grad = g.clone()
        opt.step()
        ops.step_inc(step)
        ops.adamw(p, (g * 4).cuda(), m, v, step, lr=1e-2, weight_decay=0.1, grad_scale=0.25)
    torch.cuda.synchronize()
    torch.testing.assert_close(p.cpu(), pt.detach(), rtol=1e-5, atol=1e-6)
    X = torch.randn(300, 520, generator=gen).to(torch.bfloat16)
    cs = torch.empty(520, device="cuda")
    ops.colsum(X.cuda(), cs)
    torch.testing.assert_close(cs.cpu(), X.float().sum(0), rtol=1e-4, atol=1e-3)
    q = torch.tensor([0.5, 0.005, -0.003, 0.0, -2.0, 0.01], device="cuda")
    ops.truncate_(q, 1e-2)
    assert q.cpu().tolist() == pytest.approx([0.5, 0.01, -0.01, 0.0, -2.0, 0.01])
    Xt = torch.empty(520, 300, device="cuda", dtype=torch.bfloat16)
    ops.transpose(X.cuda(), Xt)
    assert torch.equal(Xt.cpu(), X.t())


# ------------------------------------------------------------------ activation statistics (LET initialisation inputs)
@pytest.mark.parametrize("rows,C,pad", [(48, 200, 0), (2048, 4096, 0), (333, 1000, 8), (7, 8, 0), (129, 4100, 4)])
def test_act_stats_bit_exact(ops, rows, C, pad):
    """oq_act_stats / oq_act_stats_fold against the oracle (generate_act_scale_shift.py arithmetic), bit-exact;
    ragged channel counts and strided rows included."""
    from oracle import oq_oracle as orc
    g = torch.Generator().manual_seed(rows * 7 + C)
    n = 5
    xs = (torch.randn(n, rows, C + pad, generator=g) * (1 + 3 * torch.rand(C + pad, generator=g)) +
          torch.randn(C + pad, generator=g)).bfloat16()
    xd = xs.cuda()
    am = torch.empty(n, C, device="cuda")
    md = torch.empty(n, C, device="cuda")
    for i in range(n):
        ops.act_stats(xd[i, :, :C], am[i], md[i])
    per = [orc.act_stats_sample(xs[i, :, :C].float()) for i in range(n)]
    for i in range(n):
        assert torch.equal(am[i].cpu(), per[i][0]) and torch.equal(md[i].cpu(), per[i][1])
    scale, shift = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.act_stats_fold(am, md, scale, shift, first=True)
    rs, rsh = orc.act_stats_fold([p[0] for p in per], [p[1] for p in per])
    assert torch.equal(scale.cpu(), rs) and torch.equal(shift.cpu(), rsh)
    # continuing a running state in two calls gives the same result as one call
    s2, h2 = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.act_stats_fold(am[:2], md[:2], s2, h2, first=True)
    ops.act_stats_fold(am[2:].contiguous(), md[2:].contiguous(), s2, h2, first=False)
    assert torch.equal(s2, scale) and torch.equal(h2, shift)


def test_act_stats_matches_reference_fixture(ops):
    from golden_util import load
    z = load("act_stats.npz")
    xs = torch.from_numpy(z["in/x0"]).squeeze(1).bfloat16().cuda()      # bf16-representable by construction
    n, rows, C = xs.shape
    am, md = torch.empty(n, C, device="cuda"), torch.empty(n, C, device="cuda")
    for i in range(n):
        ops.act_stats(xs[i], am[i], md[i])
    scale, shift = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    ops.act_stats_fold(am, md, scale, shift)
    assert np.array_equal(scale.cpu().numpy(), z["out/scale/0"])
    assert np.array_equal(shift.cpu().numpy(), z["out/shift/0"])


# ------------------------------------------------------------------ real-quant export (GPTQ wire format)
@pytest.mark.parametrize("bits,group,N,K", [(4, 128, 64, 256), (3, 128, 96, 384), (2, 64, 64, 256), (8, 0, 32, 128),
                                            (3, 0, 64, 160), (4, 0, 4096, 4096), (3, 512, 128, 1024)])
def test_gptq_pack_bit_exact(ops, bits, group, N, K):
    """oq_gptq_pack against the numpy restatement of AutoGPTQ's QuantLinear.pack (oracle/gptq_pack.py), bit-exact, and
    the round trip packed -> codes / zero points -> dequantised weights."""
    from oracle import gptq_pack as gp
    g = torch.Generator().manual_seed(bits * 100 + N)
    W = (torch.randn(N, K, generator=g) * 0.02).cuda()
    G = K // group if group else 1
    up = torch.full((N * G,), 3.0, device="cuda")
    wq = torch.empty(N, K, device="cuda")
    sc, zp = torch.empty(N * G, device="cuda"), torch.empty(N * G, device="cuda")
    codes = torch.empty(N, K, device="cuda", dtype=torch.uint8)
    ops.weight_quant_fwd(W, n_bits=bits, group_size=group, up=up, low=up.clone(), wq_f32=wq, codes=codes, scale_out=sc,
                         zp_out=zp, recip_div=False)
    qw, qz, sh, gi = ops.gptq_pack(wq, sc, zp, n_bits=bits, group_size=group)
    ref = gp.pack(wq.cpu().numpy(), sc.cpu().numpy().reshape(N, G), zp.cpu().numpy().reshape(N, G), bits, group)
    assert np.array_equal(qw.cpu().numpy(), ref["qweight"])
    assert np.array_equal(qz.cpu().numpy(), ref["qzeros"])
    assert np.array_equal(sh.cpu().numpy(), ref["scales"])
    assert np.array_equal(gi.cpu().numpy(), ref["g_idx"])
    packed = dict(qweight=qw.cpu().numpy(), qzeros=qz.cpu().numpy(), scales=sh.cpu().numpy(), g_idx=gi.cpu().numpy())
    c2, z2 = gp.unpack(packed, bits, N, K)
    assert np.array_equal(c2, codes.cpu().numpy().astype(np.int64))            # the calibration's integer codes
    assert np.array_equal(z2, zp.cpu().numpy().reshape(N, G).astype(np.int64) & ((1 << bits) - 1))
    dq = gp.dequant(packed, bits, N, K)
    ok = (zp.cpu().numpy().reshape(N, G) >= 0).all()
    if ok:
        assert np.abs(dq - wq.cpu().numpy()).max() <= 6e-4 * np.abs(wq.cpu().numpy()).max() + 1e-8   # fp16 scales


def test_gemm_s8_exact(ops):
    """oq_gemm_s8: the integer K sum is exact, the scaled fp32 result equals the same expression evaluated by torch bit for
    bit; zero-point correction terms, bias and residual, ragged M / N / K."""
    gen = torch.Generator().manual_seed(17)
    for (M, N, K) in [(256, 512, 4096), (200, 300, 1008), (2048, 4096, 11008), (130, 96, 64)]:
        A = torch.randint(-63, 64, (M, K), generator=gen, dtype=torch.int8).cuda()
        B = torch.randint(-63, 64, (N, K), generator=gen, dtype=torch.int8).cuda()
        rs = (torch.rand(M, generator=gen) + 0.5).cuda()
        cs = (torch.rand(N, generator=gen) * 0.01 + 0.001).cuda()
        acc = (A.double() @ B.double().t())                      # exact: |acc| < 2^53
        assert float(acc.abs().max()) < 2 ** 24
        ref = acc.float() * rs[:, None] * cs[None, :]
        C = torch.empty(M, N, device="cuda")
        ops.gemm_s8(A, B, C, rs, cs)
        torch.cuda.synchronize()
        assert torch.equal(C, ref), (M, N, K, float((C - ref).abs().max()))
        # correction terms + bias + residual, bf16 output
        rt, rb = torch.randn(M, generator=gen).cuda(), (torch.randn(M, generator=gen) * 0.1).cuda()
        ct, cb = torch.randn(N, generator=gen).cuda(), (torch.randn(N, generator=gen) * 0.1).cuda()
        bias = torch.randn(N, generator=gen).cuda()
        res = torch.randn(M, N, generator=gen).to(torch.bfloat16).cuda()
        Cb = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm_s8(A, B, Cb, rs, cs, corr=(rt, rb, ct, cb), bias=bias, residual=res)
        torch.cuda.synchronize()
        full = ref + rt[:, None] * cb[None, :] + rb[:, None] * ct[None, :] + K * rb[:, None] * cb[None, :] + bias + res.float()
        torch.testing.assert_close(Cb.float(), full, rtol=1e-2, atol=1e-2 * float(full.abs().max()))


# ---------------------------------------------------------------- exact-integer operands (s8 side outputs)
def _s8_reconstruct(o):
    """(code - z0) * scale + beta == (code - zp) * scale: the dequantised row the integer operand stands for."""
    return o.codes.float() * o.scale[:, None] + o.beta[:, None]


@pytest.mark.parametrize("n_bits", [4, 6, 8])
def test_s8_side_output_of_the_activation_quantisers(ops, n_bits):
    """oq_act_quant_fwd / oq_norm_act_quant_fwd / oq_swiglu_act_quant_fwd with an s8 operand: the bf16 fake-quant output is
    unchanged (bit for bit equal to the call without the side output), the integer operand reconstructs the same
    dequantised values, t is the scaled code sum, and rows whose zero point leaves the int8-representable range (all-positive
    / all-negative rows) carry the remainder in beta."""
    gen = torch.Generator(device="cuda").manual_seed(5)
    rows, C = 96, 512
    x = torch.randn(rows, C, device="cuda", generator=gen)
    x[3] = x[3].abs() + 2.0          # zero point far below 0
    x[7] = -x[7].abs() - 1.0         # zero point above qmax
    xb = x.to(torch.bfloat16)
    y0, y1 = torch.empty_like(xb), torch.empty_like(xb)
    o = ops.S8Operand(rows, C, "cuda")
    ops.act_quant_fwd(xb, y0, n_bits=n_bits)
    ops.act_quant_fwd(xb, y1, n_bits=n_bits, s8=o)
    torch.cuda.synchronize()
    assert torch.equal(y0, y1)
    deq = orc.fake_quant(xb.float(), n_bits, recip_div=True)     # fp32 dequantised values (before the bf16 store)
    torch.testing.assert_close(_s8_reconstruct(o), deq, rtol=2e-6, atol=1e-6)
    torch.testing.assert_close(o.t, o.scale * o.codes.float().sum(1), rtol=1e-6, atol=1e-6)
    assert float(o.beta[3].abs()) > 0 and float(o.beta[7].abs()) > 0
    if n_bits < 8:
        assert int((o.beta == 0).sum()) >= rows - 2            # rows that straddle zero need no correction
    assert int(o.codes.int().abs().max()) <= 128
    # norm + quantiser
    w = torch.rand(C, device="cuda", generator=gen) + 0.5
    yn, q0, q1 = torch.empty_like(xb), torch.empty_like(xb), torch.empty_like(xb)
    o2 = ops.S8Operand(rows, C, "cuda")
    ops.norm_act_quant_fwd(xb, w, None, 1e-6, 0, yn, q0, n_bits=n_bits)
    ops.norm_act_quant_fwd(xb, w, None, 1e-6, 0, yn, q1, n_bits=n_bits, s8=o2)
    torch.cuda.synchronize()
    assert torch.equal(q0, q1)
    torch.testing.assert_close(_s8_reconstruct(o2), orc.fake_quant(yn.float(), n_bits, recip_div=True), rtol=2e-6, atol=1e-6)
    torch.testing.assert_close(o2.t, o2.scale * o2.codes.float().sum(1), rtol=1e-6, atol=1e-6)
    # SwiGLU + quantiser
    F_ = 768
    gu = torch.randn(rows, 2 * F_, device="cuda", generator=gen).to(torch.bfloat16)
    m0, m1 = torch.empty(rows, F_, device="cuda", dtype=torch.bfloat16), torch.empty(rows, F_, device="cuda", dtype=torch.bfloat16)
    o3 = ops.S8Operand(rows, F_, "cuda")
    ops.swiglu_act_quant_fwd(gu, m0, n_bits=n_bits)
    ops.swiglu_act_quant_fwd(gu, m1, n_bits=n_bits, s8=o3)
    torch.cuda.synchronize()
    assert torch.equal(m0, m1)
    torch.testing.assert_close(_s8_reconstruct(o3).to(torch.bfloat16).float(), m1.float(), rtol=1e-2, atol=1e-6)
    torch.testing.assert_close(o3.t, o3.scale * o3.codes.float().sum(1), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("n_bits,let", [(4, True), (6, False), (8, True)])
def test_s8_side_output_of_the_weight_quantiser(ops, n_bits, let):
    gen = torch.Generator(device="cuda").manual_seed(9)
    N, K = 320, 1024
    W = (torch.randn(N, K, device="cuda", generator=gen) * 0.05).half().float()
    W[5] = W[5].abs() + 0.1           # a channel that does not straddle zero (zero point below 0 even at 4 bits)
    up = torch.full((N,), 3.0, device="cuda")
    kw = dict(col_scale=torch.rand(K, device="cuda", generator=gen) + 0.5,
              row_scale=torch.rand(N, device="cuda", generator=gen) + 0.5, row_op=2) if let else {}
    ref = run_wq(ops, W, n_bits, False, 0, up, up, True, **kw)
    o = ops.S8Operand(N, K, "cuda")
    out = dict(wq_f32=torch.empty(N, K, device="cuda"), zp_out=torch.empty(N, device="cuda"))
    ops.weight_quant_fwd(W, n_bits=n_bits, up=up, low=up, recip_div=True, s8=o, **out, **kw)
    torch.cuda.synchronize()
    assert torch.equal(out["wq_f32"], ref["wq_f32"]) and torch.equal(o.scale, ref["scale_out"]) and torch.equal(out["zp_out"], ref["zp_out"])
    torch.testing.assert_close(_s8_reconstruct(o), ref["wq_f32"], rtol=2e-6, atol=1e-7)
    torch.testing.assert_close(o.t, o.scale * o.codes.float().sum(1), rtol=1e-6, atol=1e-7)
    assert float(o.beta[5].abs()) > 0
    z0 = ref["zp_out"].clamp(max(0, 2 ** n_bits - 1 - 127), min(2 ** n_bits - 1, 128))
    assert torch.equal(o.codes.float(), ref["codes"].float() - z0[:, None])


def test_linear_s8_equals_the_linear_of_the_fake_quantised_operands(ops):
    """QuantLinear forward from the integer operands (quantiser side outputs -> oq_gemm_s8) against F.linear of the fp32
    dequantised operands, LLaMA-7B down_proj-like K, including rows / channels with out-of-range zero points."""
    gen = torch.Generator(device="cuda").manual_seed(11)
    M, N, K = 384, 512, 11008
    x = torch.randn(M, K, device="cuda", generator=gen)
    x[1] = x[1].abs() + 0.5
    W = (torch.randn(N, K, device="cuda", generator=gen) * 0.02).half().float()
    W[2] = -W[2].abs() - 0.001
    xb = x.to(torch.bfloat16)
    for bits in (4, 6):
        xq, wq = torch.empty_like(xb), torch.empty(N, K, device="cuda")
        xo, wo = ops.S8Operand(M, K, "cuda"), ops.S8Operand(N, K, "cuda")
        ops.act_quant_fwd(xb, xq, n_bits=bits, s8=xo)
        ops.weight_quant_fwd(W, n_bits=bits, recip_div=True, wq_f32=wq, s8=wo)
        bias = torch.randn(N, device="cuda", generator=gen)
        res = torch.randn(M, N, device="cuda", generator=gen).to(torch.bfloat16)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.linear_s8_fwd(xo, wo, out, bias=bias, residual=res, alpha=0.5)
        torch.cuda.synchronize()
        xdeq = orc.fake_quant(xb.float(), bits, recip_div=True)
        want = 0.5 * (xdeq.double() @ wq.double().t()).float() + bias + res.float()
        torch.testing.assert_close(out.float(), want, rtol=8e-3, atol=8e-3 * float(want.abs().mean()))
        # tighter: fp32 output, no bf16 rounding of the result
        out32 = torch.empty(M, N, device="cuda")
        ops.linear_s8_fwd(xo, wo, out32)
        torch.cuda.synchronize()
        want32 = (xdeq.double() @ wq.double().t()).float()
        torch.testing.assert_close(out32, want32, rtol=1e-4, atol=1e-4 * float(want32.abs().mean()))


def test_engine_s8_forward_agrees_with_the_bf16_forward(monkeypatch):
    """The same W4A4 block stepped with the exact-integer forward linears and with the bf16 ones: first-step loss within
    the bf16 operand-rounding band, gradients aligned."""
    from omniquant_b200.engine import BlockEngine, QuantConfig
    g = torch.Generator().manual_seed(0)
    H, F_, heads, T = 512, 1024, 4, 128
    r = lambda *s, sc=0.03: (torch.randn(*s, generator=g) * sc).half().float()
    W = {"q": r(H, H), "k": r(H, H), "v": r(H, H), "o": r(H, H), "gate": r(F_, H), "up": r(F_, H), "down": r(H, F_),
         "ln1": 1 + r(H, sc=0.1), "ln2": 1 + r(H, sc=0.1)}
    x = torch.randn(T, H, generator=g).cuda().to(torch.bfloat16)
    tgt = torch.roll(x, 1, 0)
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("OQ_S8", flag)
        eng = BlockEngine("llama", W, n_heads=heads, seq_len=T, eps=1e-6,
                          qcfg=QuantConfig(wbits=4, abits=4, group_size=0, lwc=True, let=True, recip_div=True))
        assert eng.s8 == (flag == "1")
        loss = float(eng.train_step(x, tgt).clone())
        torch.cuda.synchronize()
        res.append((loss, eng.grads.clone()))
    assert abs(res[0][0] - res[1][0]) / res[1][0] < 5e-3, (res[0][0], res[1][0])
    a, b = res[0][1].double(), res[1][1].double()
    assert float((a @ b) / (a.norm() * b.norm())) > 0.9      # 4-bit codes downstream flip on 2^-9 operand differences (0.968 measured)


# ---------------------------------------------------------------- fused attention (QK^T -> softmax -> PV, recompute backward)
def _attn_reference(q, k, v, alpha):
    """fp32 torch evaluation of QuantMatMul attention on [heads, T, d] operands with the engine's rounding points
    (scores to fp16, probabilities to bf16, straight-through gradients), causal."""
    T = q.shape[1]
    s = torch.matmul(q, k.transpose(1, 2)) * alpha
    s = s + (s.half().float() - s).detach()
    mask = torch.triu(torch.ones(T, T, dtype=torch.bool, device=q.device), diagonal=1)
    s = s.masked_fill(mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = p + (p.to(torch.bfloat16).float() - p).detach()
    return torch.matmul(p, v)


@pytest.mark.parametrize("layout,heads,kv,T,batch,d", [("llama", 4, 4, 256, 1, 128), ("llama", 8, 2, 384, 2, 128),
                                                       ("falcon", 6, 2, 256, 1, 128), ("llama", 2, 2, 2048, 1, 128),
                                                       ("llama", 4, 4, 384, 2, 64), ("falcon", 8, 2, 256, 1, 64),
                                                       ("llama", 3, 3, 2048, 1, 64)])
def test_fused_attention_fwd_bwd(ops, layout, heads, kv, T, batch, d):
    """oq_attn_fwd / oq_attn_bwd against the fp32 torch evaluation of the same expression (MHA, grouped-query heads in
    the LLaMA [q | k | v] and the Falcon per-group layouts, several samples, head dimensions 128 and 64 -- OPT-125m and
    Falcon-180B have 64)."""
    rep = heads // kv
    gen = torch.Generator(device="cuda").manual_seed(heads * 131 + T + d)
    H, Hkv = heads * d, kv * d
    QW = H + 2 * Hkv
    M = batch * T
    qkv = (torch.randn(M, QW, device="cuda", generator=gen) * 1.5).to(torch.bfloat16)
    dO = torch.randn(M, H, device="cuda", generator=gen).to(torch.bfloat16)
    if layout == "falcon":
        gw = (rep + 2) * d
        cols = dict(cq=0, gq=gw, ck=rep * d, cv=(rep + 1) * d, gk=gw)
        view = qkv.view(M, kv, rep + 2, d)
        qf, kf, vf = view[:, :, :rep].reshape(M, heads, d), view[:, :, rep], view[:, :, rep + 1]
    else:
        cols = dict(cq=0, gq=rep * d, ck=H, cv=H + Hkv, gk=d)
        qf, kf, vf = qkv[:, :H].view(M, heads, d), qkv[:, H:H + Hkv].view(M, kv, d), qkv[:, H + Hkv:].view(M, kv, d)
    alpha = d ** -0.5
    O = torch.empty(M, H, device="cuda", dtype=torch.bfloat16)
    rmax, rinv = torch.empty(batch, heads, T, device="cuda"), torch.empty(batch, heads, T, device="cuda")
    kw = dict(batch=batch, T=T, heads=heads, kv_heads=kv, head_dim=d, layout=cols, alpha=alpha)
    ops.attn_fwd(qkv, O, rmax, rinv, **kw)
    dqkv = torch.full_like(qkv, float("nan"))
    ops.attn_bwd(qkv, O, dO, rmax, rinv, torch.empty(batch, heads, T, device="cuda"), dqkv, **kw)
    torch.cuda.synchronize()
    assert bool(torch.isfinite(O.float()).all()) and bool(torch.isfinite(dqkv.float()).all())   # every element written
    # fp32 reference with autograd, sample by sample
    want_O = torch.empty(M, H, device="cuda")
    want_g = torch.zeros(M, QW, device="cuda")
    for s in range(batch):
        r = slice(s * T, (s + 1) * T)
        q = qf[r].float().transpose(0, 1).contiguous().requires_grad_(True)           # [heads, T, d]
        k = kf[r].float().transpose(0, 1).contiguous().requires_grad_(True)           # [kv, T, d]
        v = vf[r].float().transpose(0, 1).contiguous().requires_grad_(True)
        o = _attn_reference(q, k.repeat_interleave(rep, 0), v.repeat_interleave(rep, 0), alpha)
        want_O[r] = o.transpose(0, 1).reshape(T, H)
        o.backward(dO[r].float().view(T, heads, d).transpose(0, 1))
        gq, gk, gv = (t.grad.transpose(0, 1) for t in (q, k, v))
        if layout == "falcon":
            gview = want_g[r].view(T, kv, rep + 2, d)
            gview[:, :, :rep] = gq.reshape(T, kv, rep, d)
            gview[:, :, rep], gview[:, :, rep + 1] = gk, gv
        else:
            want_g[r, :H], want_g[r, H:H + Hkv], want_g[r, H + Hkv:] = gq.reshape(T, H), gk.reshape(T, Hkv), gv.reshape(T, Hkv)
    rel = lambda a, b: float((a.float() - b).norm() / b.norm())
    print(f"fused attention {layout} h{heads}/{kv} T{T} b{batch}: O rel {rel(O, want_O):.2e}  grad rel {rel(dqkv, want_g):.2e}")
    assert rel(O, want_O) < 4e-3
    assert rel(dqkv, want_g) < 1.2e-2
    # softmax statistics
    s0 = (torch.matmul(qf[:T].float().transpose(0, 1), kf[:T].float().transpose(0, 1).repeat_interleave(rep, 0).transpose(1, 2)) * alpha).half().float()
    s0 = s0.masked_fill(torch.triu(torch.ones(T, T, dtype=torch.bool, device="cuda"), diagonal=1), float("-inf"))
    torch.testing.assert_close(rmax[0], s0.amax(-1), rtol=0, atol=2e-2)


def test_engine_fused_attention_agrees_with_the_unfused_chain(monkeypatch):
    """One optimiser step of the same W4A4 block with the fused attention kernels and with batched GEMMs + softmax kernels."""
    from omniquant_b200.engine import BlockEngine, QuantConfig
    g = torch.Generator().manual_seed(0)
    H, F_, heads, T = 512, 1024, 4, 256
    r = lambda *s, sc=0.03: (torch.randn(*s, generator=g) * sc).half().float()
    W = {"q": r(H, H), "k": r(H, H), "v": r(H, H), "o": r(H, H), "gate": r(F_, H), "up": r(F_, H), "down": r(H, F_),
         "ln1": 1 + r(H, sc=0.1), "ln2": 1 + r(H, sc=0.1)}
    x = torch.randn(T, H, generator=g).cuda().to(torch.bfloat16)
    tgt = torch.roll(x, 1, 0)
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("OQ_FUSED_ATTN", flag)
        eng = BlockEngine("llama", W, n_heads=heads, seq_len=T, eps=1e-6,
                          qcfg=QuantConfig(wbits=4, abits=16, group_size=0, lwc=True, let=True, recip_div=True))
        assert eng.fused_attn == (flag == "1")
        loss = float(eng.train_step(x, tgt).clone())
        torch.cuda.synchronize()
        res.append((loss, eng.grads.clone(), eng.bufs["O"].clone()))
    assert float((res[0][2].float() - res[1][2].float()).norm() / res[1][2].float().norm()) < 2e-3
    assert abs(res[0][0] - res[1][0]) / res[1][0] < 2e-3, (res[0][0], res[1][0])
    a, b = res[0][1].double(), res[1][1].double()
    assert float((a - b).norm() / b.norm()) < 3e-2


def test_weight_quant_row_segments_equal_separate_launches(ops):
    """q | k | v in ONE launch (row segment table: shared column scale / shift, per-weight row vector and operation) against
    three separate launches: forward outputs bit for bit, backward reductions to atomic-order round-off."""
    gen = torch.Generator(device="cuda").manual_seed(13)
    H, K = 384, 512
    W = (torch.randn(3 * H, K, device="cuda", generator=gen) * 0.05).half().float()
    col = torch.rand(K, device="cuda", generator=gen) + 0.5
    shift = torch.randn(K, device="cuda", generator=gen) * 0.1
    rq, rv = torch.rand(H, device="cuda", generator=gen) + 0.5, torch.rand(H, device="cuda", generator=gen) + 0.5
    up = torch.randn(3 * H, device="cuda", generator=gen) + 3
    low = torch.randn(3 * H, device="cuda", generator=gen) + 3
    g = (torch.randn(3 * H, K, device="cuda", generator=gen) * 0.01).to(torch.bfloat16)
    e = torch.randn(3 * H, device="cuda", generator=gen)
    segs = [(H, rq, 2), (2 * H, rq, 1), (3 * H, rv, 2)]
    z = lambda *s, dt=torch.float32: torch.zeros(*s, device="cuda", dtype=dt)

    def fwd(Wp, upp, lowp, **kw):
        n = Wp.shape[0]
        out = dict(wq_f32=z(n, K), wq_bf16=z(n, K, dt=torch.bfloat16), scale_out=z(n), zp_out=z(n), xmin_out=z(n), xmax_out=z(n),
                   matvec_out=z(n))
        ops.weight_quant_fwd(Wp, n_bits=4, up=upp, low=lowp, col_scale=col, shift=shift, recip_div=True, **out, **kw)
        return out

    def bwd(Wp, gp, upp, lowp, st, ep, **kw):
        n = Wp.shape[0]
        out = dict(d_up=z(n), d_low=z(n), dcol_acc=z(K), drow=z(n), dshift_acc=z(K))
        ops.weight_quant_bwd(Wp, gp, st["xmin_out"], st["xmax_out"], n_bits=4, up=upp, low=lowp, col_scale=col, recip_div=True,
                             e_row=ep, scale=st["scale_out"], zp=st["zp_out"], **out, **kw)
        return out

    one = fwd(W, up, low, segments=segs)
    oneb = bwd(W, g, up, low, one, e, segments=segs)
    parts, partsb = [], []
    for i, (row, op) in enumerate(((rq, 2), (rq, 1), (rv, 2))):
        r = slice(i * H, (i + 1) * H)
        parts.append(fwd(W[r].contiguous(), up[r].contiguous(), low[r].contiguous(), row_scale=row, row_op=op))
        partsb.append(bwd(W[r].contiguous(), g[r].contiguous(), up[r].contiguous(), low[r].contiguous(), parts[-1], e[r].contiguous(),
                          row_scale=row, row_op=op))
    torch.cuda.synchronize()
    for k in ("wq_f32", "wq_bf16", "scale_out", "zp_out", "xmin_out", "xmax_out"):
        assert torch.equal(one[k], torch.cat([p_[k] for p_ in parts], 0)), k
    torch.testing.assert_close(one["matvec_out"], torch.cat([p_["matvec_out"] for p_ in parts]), rtol=1e-5, atol=1e-6)
    for k in ("d_up", "d_low", "drow"):
        torch.testing.assert_close(oneb[k], torch.cat([p_[k] for p_ in partsb]), rtol=2e-4, atol=1e-6)
    for k in ("dcol_acc", "dshift_acc"):
        torch.testing.assert_close(oneb[k], sum(p_[k] for p_ in partsb), rtol=2e-4, atol=1e-5)


# ---------------------------------------------------------------- K1: weight fake-quant fused into the GEMM
@pytest.mark.parametrize("N,K,grp,bits", [(256, 512, 0, 4), (384, 1024, 128, 3), (200, 256, 64, 2), (4096, 4096, 0, 4)])
def test_gemm_fq_operand_is_bit_exact(ops, N, K, grp, bits):
    """oq_gemm_fq with A = identity returns the operand the tensor cores consumed: it must equal the bf16 output of the
    stand-alone weight quantiser bit for bit (same scale / zero point, same exact-rounding guard)."""
    gen = torch.Generator(device="cuda").manual_seed(N + K)
    W = (torch.randn(N, K, device="cuda", generator=gen) * 0.03).half().float()
    G = K // grp if grp else 1
    up = torch.randn(N * G, device="cuda", generator=gen) + 3
    low = torch.randn(N * G, device="cuda", generator=gen) + 3
    ref = run_wq(ops, W, bits, False, grp, up, low, True)
    scale, zp = ops.lwc_qparams(ref["xmin_out"], ref["xmax_out"], up, low, n_bits=bits)
    assert torch.equal(scale, ref["scale_out"]) and torch.equal(zp, ref["zp_out"])
    eye = torch.eye(K, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(K, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm_fq(eye, W, scale, zp, out, n_bits=bits, group_size=grp)
    torch.cuda.synchronize()
    assert torch.equal(out.t().contiguous(), ref["wq_bf16"]), float((out.t().float() - ref["wq_bf16"].float()).abs().max())


def test_gemm_fq_matches_prepass_plus_gemm(ops):
    """The fused kernel against the path the engine uses (oq_weight_quant_fwd -> oq_gemm_bf16), ragged M, bias + residual."""
    gen = torch.Generator(device="cuda").manual_seed(3)
    M, N, K = 700, 640, 1024
    x = torch.randn(M, K, device="cuda", generator=gen).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda", generator=gen) * 0.03).half().float()
    up = torch.randn(N, device="cuda", generator=gen) + 3
    ref = run_wq(ops, W, 4, False, 0, up, up, True)
    bias = torch.randn(N, device="cuda", generator=gen)
    res = torch.randn(M, N, device="cuda", generator=gen).to(torch.bfloat16)
    want = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.linear_fwd(x, ref["wq_bf16"], want, bias=bias, residual=res)
    got = torch.empty_like(want)
    ops.gemm_fq(x, W, ref["scale_out"], ref["zp_out"], got, n_bits=4, bias=bias, residual=res)
    torch.cuda.synchronize()
    torch.testing.assert_close(got.float(), want.float(), rtol=1e-2, atol=1e-2)
    assert float((got.float() - want.float()).abs().max()) <= 2.0 ** -6 * float(want.float().abs().max())


@pytest.mark.parametrize("N,K,let", [(64, 4096, "col_row_mul"), (48, 4096, "col_row_div"), (40, 4096, "col"), (33, 11008, None),
                                     (96, 768, "col_row_mul"), (64, 3072, None), (24, 5120, "col_row_div"), (36, 4096, "segments")])
@pytest.mark.parametrize("n_bits,symmetric", [(4, False), (6, False), (8, False), (3, True), (2, False)])
def test_weight_quant_calibration_variant_equals_general_kernel(ops, N, K, let, n_bits, symmetric):
    """The calibration-loop forward (bf16 output [+ s8 codes] only: clamp-then-round with a row-constant exactness guard)
    must produce the same bits as the general kernel, which the tests above pin to the oracle / ATen.  Rows are built to
    hit the guard: weights at exact multiples and half-multiples of the row's scale, constant rows, rows that do not
    straddle zero (zero points outside [0, qmax]), tiny and huge rows, NaN and Inf elements."""
    g = torch.Generator(device="cuda").manual_seed(N * 7 + K + n_bits)
    W = torch.randn(N, K, device="cuda", generator=g) * 0.05
    W[1] = torch.round(W[1] * 64) / 64                      # many exact ties on a dyadic grid
    W[2] = 0.37                                             # constant row: scale 0 -> NaN path
    W[3] = W[3].abs() + 0.5                                 # all positive: negative zero point
    W[4] = -W[4].abs() - 3.0                                # all negative: zero point above qmax
    W[5] = W[5] * 1e-30
    W[6] = W[6] * 1e30
    W[7, 5] = float("nan")
    W[8, 9] = float("inf")
    W[9] = torch.linspace(-1, 1, K, device="cuda")          # evenly spaced: hits k + 0.5 quotients
    W[10] = (torch.arange(K, device="cuda") % 31 - 15).float() * 0.125 + 0.0625
    up = torch.randn(N, device="cuda", generator=g) * 2 + 2
    low = torch.randn(N, device="cuda", generator=g) * 2 + 2
    kw = dict(n_bits=n_bits, symmetric=symmetric, up=up, low=low)
    if let in ("col_row_mul", "col_row_div", "col"):
        kw["col_scale"] = torch.rand(K, device="cuda", generator=g) + 0.5
        kw["shift"] = torch.randn(K, device="cuda", generator=g) * 0.1
        if let != "col":
            kw["row_scale"] = torch.rand(N, device="cuda", generator=g) + 0.5
            kw["row_op"] = 1 if let == "col_row_mul" else 2
    elif let == "segments":
        kw["col_scale"] = torch.rand(K, device="cuda", generator=g) + 0.5
        kw["shift"] = torch.randn(K, device="cuda", generator=g) * 0.1
        r = torch.rand(N, device="cuda", generator=g) + 0.5
        kw["segments"] = [(12, r[:12], 2), (24, r[12:24], 1), (N, None, 0)]
    for recip in (True, False):
        outs = []
        for general in (True, False):
            for s8 in (False, True):
                q = torch.full((N, K), 7.0, device="cuda", dtype=torch.bfloat16)
                sc, zp, xmn, xmx = (torch.zeros(N, device="cuda") for _ in range(4))
                mv = torch.zeros(N, device="cuda") if "shift" in kw else None
                op = ops.S8Operand(N, K, "cuda") if s8 else None
                extra = dict(wq_f32=torch.empty(N, K, device="cuda")) if general else {}     # any extra output selects the general kernel
                ops.weight_quant_fwd(W, recip_div=recip, wq_bf16=q, scale_out=None if s8 else sc, zp_out=zp, xmin_out=xmn,
                                     xmax_out=xmx, matvec_out=mv, s8=op, **kw, **extra)
                torch.cuda.synchronize()
                outs.append((general, s8, q.view(torch.int16), (op.scale if s8 else sc).view(torch.int32).clone(),
                             zp.view(torch.int32), xmn, xmx, mv, op))
        ref, ref8 = outs[0], outs[1]
        for o in outs[2:]:
            base = ref8 if (o[1] and ref8 is not None) else ref
            rows_ok = torch.ones(N, dtype=torch.bool, device="cuda")
            rows_ok[7] = False      # NaN element: the general kernel ignores it in the extrema, the lean one poisons the row (like torch.amax)
            assert torch.equal(o[2][rows_ok], base[2][rows_ok]), (let, n_bits, recip, o[1])
            assert torch.equal(o[3][rows_ok], base[3][rows_ok]) and torch.equal(o[4][rows_ok], base[4][rows_ok])
            assert torch.equal(o[5][rows_ok], base[5][rows_ok]) and torch.equal(o[6][rows_ok], base[6][rows_ok])
            if o[7] is not None:
                fin_rows = rows_ok & torch.isfinite(W).all(1)
                bound = (W[fin_rows].abs().double() @ kw["shift"].abs().double()).float() * 1e-5 + 1e-30     # fp32 summation order
                assert bool(((o[7][fin_rows] - base[7][fin_rows]).abs() <= bound).all())
            assert bool(torch.isnan(o[2][7].view(torch.bfloat16).float()).all())     # reference semantics for a NaN weight
            if o[1] and ref8 is not None:
                fin = rows_ok & torch.isfinite(base[8].beta)
                assert torch.equal(o[8].codes[fin], base[8].codes[fin])
                assert torch.equal(o[8].t[fin].view(torch.int32), base[8].t[fin].view(torch.int32))
                assert torch.equal(torch.isfinite(o[8].beta), torch.isfinite(base[8].beta) & rows_ok | (torch.isfinite(o[8].beta) & ~rows_ok))
                assert torch.equal(o[8].beta[fin].view(torch.int32), base[8].beta[fin].view(torch.int32))
