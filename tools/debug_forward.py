"""Stage-by-stage comparison of the engine's quantised forward with the oracle run in fp32 on the GPU."""
import math
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from golden_util import BlockFixture  # noqa: E402
from oracle import oq_oracle as orc  # noqa: E402
from test_engine_gpu import build_engine  # noqa: E402


def rel(a, b):
    a, b = a.float().cuda(), b.float().cuda()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main(fname):
    fx = BlockFixture(fname)
    eng = build_engine(fx, recip_div=True)
    cfg = orc.QCfg(fx.wbits, fx.abits, fx.group_size, False, fx.lwc, fx.let)
    Wt = {k: v.cuda() for k, v in fx.weights().items()}
    P = {k: v.cuda() for k, v in fx.init_params().items()}
    x = fx.inps[:fx.batch_size].cuda()
    xb = x.to(torch.bfloat16)
    T_, H, h = fx.T, fx.H, fx.heads
    d = H // h
    eng.prepare()
    out = eng.forward(xb.view(-1, H), eng.tmp, act_quant=True).clone()
    sv = eng.saved
    Tq = orc.llama_prepare(Wt, P, cfg)
    print("W qkv", rel(eng.Wq["qkv"], torch.cat([Tq["q.w"], Tq["k.w"], Tq["v.w"]])))
    print("W o", rel(eng.Wq["o"], Tq["o.w"]), "W fc1", rel(eng.Wq["fc1"], torch.cat([Tq["gate.w"], Tq["up.w"]])),
          "W fc2", rel(eng.Wq["fc2"], Tq["down.w"]))
    print("bqkv", rel(eng.tmp["bqkv"], torch.cat([Tq["q.b"], Tq["k.b"], Tq["v.b"]])), "ln1w", rel(eng.tmp["ln1w"], Tq["ln1w"]),
          "ln1b", rel(eng.tmp["ln1b"], Tq["ln1b"]))
    xf = xb.float()           # same bf16-rounded input
    a1 = orc.rms_norm(xf, Tq["ln1w"], Tq["ln1b"], 1e-6)
    print("a1", rel(sv["a1"].view_as(a1), a1))
    a1q = orc.aq(a1, cfg)
    print("a1q", rel(sv["a1q"].view_as(a1q), a1q))
    b = x.shape[0]
    q = F.linear(a1q, Tq["q.w"], Tq["q.b"]); k = F.linear(a1q, Tq["k.w"], Tq["k.b"]); v = F.linear(a1q, Tq["v.w"], Tq["v.b"])
    print("qkv", rel(eng.qkv.view(b, T_, 3 * H), torch.cat([q, k, v], -1)))
    cos, sin = orc.rope_tables(d, T_)
    cos, sin = cos.cuda(), sin.cuda()
    qh = q.view(b, T_, h, d).transpose(1, 2); kh = k.view(b, T_, h, d).transpose(1, 2); vh = v.view(b, T_, h, d).transpose(1, 2)
    qh, kh = orc.apply_rope(qh, kh, cos, sin)
    qh, kh, vh = orc.aq(qh, cfg), orc.aq(kh, cfg), orc.aq(vh, cfg)
    eq = eng.qkvq.view(b, T_, 3, h, d)
    print("qhat", rel(eq[:, :, 0].transpose(1, 2), qh), "khat", rel(eq[:, :, 1].transpose(1, 2), kh), "vhat",
          rel(eq[:, :, 2].transpose(1, 2), vh))
    s = qh @ kh.transpose(2, 3) / math.sqrt(d) + orc.causal_mask(T_).cuda()
    p = torch.softmax(s, -1)
    print("P", rel(sv["P"], p))
    o = (p @ vh).transpose(1, 2).reshape(b, T_, H)
    print("O", rel(sv["O"].view_as(o), o))
    oq = orc.aq(o, cfg)
    print("Oq", rel(sv["Oq"].view_as(o), oq))
    h1 = xf + F.linear(oq, Tq["o.w"], Tq["o.b"])
    print("h1", rel(sv["h1"].view_as(h1), h1))
    a2 = orc.rms_norm(h1, Tq["ln2w"], Tq["ln2b"], 1e-6)
    a2q = orc.aq(a2, cfg)
    print("a2q", rel(sv["a2q"].view_as(a2q), a2q))
    g = F.linear(a2q, Tq["gate.w"], Tq["gate.b"]); u = F.linear(a2q, Tq["up.w"], Tq["up.b"])
    print("gu", rel(sv["f1"].view(b, T_, -1), torch.cat([g, u], -1)))
    m = F.silu(g) * u
    print("m", rel(sv["m"].view_as(m), m))
    mq = orc.aq(m, cfg)
    print("mq", rel(sv["mq"].view_as(m), mq))
    o2 = h1 + F.linear(mq, Tq["down.w"], Tq["down.b"])
    print("out", rel(out.view_as(o2), o2))
    Tfp = orc.llama_prepare(Wt, P, orc.QCfg(fx.wbits, fx.abits, fx.group_size, False, fx.lwc, False), quant=False)
    fp = orc.llama_forward(xf, Tfp, h, cos, sin, orc.causal_mask(T_).cuda(), cfg, False)
    fpe = eng.forward(xb.view(-1, H), eng.fp_weights(), act_quant=False).clone()
    print("fp_out", rel(fpe.view_as(fp), fp))
    print("loss oracle(bf16 x)", float(F.mse_loss(fp, o2)), "engine", float(F.mse_loss(fpe.float().view_as(fp), out.float().view_as(fp))),
          "engine-out vs oracle-fp", float(F.mse_loss(fp, out.float().view_as(fp))), "ref", fx.losses[0])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "llama_w4a4_lwc_let.npz")
