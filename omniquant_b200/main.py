"""Command line of the calibration run with the reference's flags (reference main.py:194-229) and the
six quantiser parameter dicts it derives (main.py:268-303).

    python -m omniquant_b200.main --model <hf path | llama-7b | opt-125m> --wbits 4 --abits 4 --lwc --let --epochs 20

Evaluation (perplexity / lm_eval, main.py:55-187) and dataset loaders (datautils.py) need checkpoints,
datasets and network access and are outside the calibration hot path: ``--synthetic`` (default when the
model path is not a local directory) calibrates a random-init model of the named architecture on
synthetic tokens, which is also what bench.py measures.
"""
import argparse
import os
import random
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from . import utils
from .quantize.omniquant import omniquant

net_choices = ["opt-125m", "opt-1.3b", "opt-2.7b", "opt-6.7b", "opt-13b", "opt-30b", "opt-66b", "llama-7b", "llama-13b",
               "llama-30b", "llama-65b", "Llama-2-7b", "Llama-2-13b", "Llama-2-70b", "Llama-2-7b-chat", "Llama-2-13b-chat"]

ARCH = {  # hidden, ffn, heads, layers
    "opt-125m": (768, 3072, 12, 12), "opt-1.3b": (2048, 8192, 32, 24), "opt-6.7b": (4096, 16384, 32, 32),
    "llama-7b": (4096, 11008, 32, 32), "Llama-2-7b": (4096, 11008, 32, 32), "llama-13b": (5120, 13824, 40, 40),
    "Llama-2-13b": (5120, 13824, 40, 40),
    "llama-tiny": (256, 512, 2, 2), "opt-tiny": (256, 1024, 4, 2), "falcon-tiny": (256, 1024, 4, 2),   # smoke-test sizes
    "falcon-180b": (14848, 59392, 232, 80), "Llama-2-70b": (8192, 28672, 64, 80),
}


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--model", type=str, help="model name of model path")
    p.add_argument("--cache_dir", default="./cache", type=str)
    p.add_argument("--output_dir", default="../log/", type=str)
    p.add_argument("--save_dir", default=None, type=str)
    p.add_argument("--resume", type=str, default=None)
    p.add_argument("--real_quant", default=False, action="store_true")
    p.add_argument("--calib_dataset", type=str, default="wikitext2", choices=["wikitext2", "ptb", "c4", "mix", "pile"])
    p.add_argument("--nsamples", type=int, default=128)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--seed", type=int, default=2)
    p.add_argument("--tasks", default="")
    p.add_argument("--eval_ppl", action="store_true")
    p.add_argument("--num_fewshot", type=int, default=0)
    p.add_argument("--wbits", type=int, default=4)
    p.add_argument("--abits", type=int, default=4)
    p.add_argument("--group_size", type=int, default=None)
    p.add_argument("--alpha", type=float, default=0.5)
    p.add_argument("--let_lr", type=float, default=5e-3)
    p.add_argument("--lwc_lr", type=float, default=1e-2)
    p.add_argument("--wd", type=float, default=0)
    p.add_argument("--epochs", type=int, default=10)
    p.add_argument("--let", default=False, action="store_true")
    p.add_argument("--lwc", default=False, action="store_true")
    p.add_argument("--aug_loss", default=False, action="store_true")
    p.add_argument("--symmetric", default=False, action="store_true")
    p.add_argument("--a_dynamic_method", type=str, default="per_token", choices=["per_token"])
    p.add_argument("--w_dynamic_method", type=str, default="per_channel", choices=["per_channel"])
    p.add_argument("--limit", type=int, default=-1)
    p.add_argument("--multigpu", action="store_true")
    p.add_argument("--deactive_amp", action="store_true")
    p.add_argument("--net", type=str, default=None)
    p.add_argument("--act-scales", type=str, default=None)
    p.add_argument("--act-shifts", type=str, default=None)
    # additions of this implementation
    p.add_argument("--synthetic", action="store_true", help="random-init model of --net and synthetic tokens")
    p.add_argument("--num_layers", type=int, default=None, help="calibrate only the first N blocks (synthetic runs)")
    p.add_argument("--online_act_stats", action="store_true",
                   help="collect the LET activation statistics in the FP pass of each block instead of loading "
                        "--act-scales / --act-shifts (generate_act_scale_shift.py); also used when those files are absent")
    p.add_argument("--seqlen", type=int, default=2048)
    return p


def quant_param_dicts(args):
    """main.py:268-303."""
    args.weight_quant_params = {"n_bits": args.wbits, "per_channel_axes": [0], "symmetric": args.symmetric,
                                "dynamic_method": args.w_dynamic_method, "group_size": args.group_size, "lwc": args.lwc}
    tok = {"n_bits": args.abits, "per_channel_axes": [], "symmetric": False, "dynamic_method": args.a_dynamic_method}
    args.act_quant_params = dict(tok)
    args.q_quant_params = dict(tok)
    args.k_quant_params = dict(tok)
    args.v_quant_params = dict(tok)
    args.p_quant_params = {"n_bits": 16, "metric": "fix0to1"}
    return args


def synthetic_model(net, num_layers=None, seed=0):
    """Random-init HF model of the named architecture holding fp16-representable weights (on CPU)."""
    from transformers import LlamaConfig, LlamaForCausalLM, OPTConfig, OPTForCausalLM
    H, F_, heads, L = ARCH[net]
    L = num_layers or L
    torch.manual_seed(seed)
    if "falcon" in net.lower():
        from transformers import FalconConfig, FalconForCausalLM
        cfg = FalconConfig(hidden_size=H, num_attention_heads=heads, num_kv_heads=8 if H >= 768 else 2, num_hidden_layers=L,
                           vocab_size=65024 if H >= 768 else 512, new_decoder_architecture=True, bias=False,
                           parallel_attn=True, ffn_hidden_size=F_)
        model = FalconForCausalLM(cfg)
    elif "opt" in net.lower():
        cfg = OPTConfig(hidden_size=H, ffn_dim=F_, num_attention_heads=heads, num_hidden_layers=L,
                        vocab_size=50272 if H >= 768 else 512, max_position_embeddings=2048, word_embed_proj_dim=H)
        model = OPTForCausalLM(cfg)
    else:
        cfg = LlamaConfig(hidden_size=H, intermediate_size=F_, num_attention_heads=heads,
                          num_key_value_heads=8 if net == "Llama-2-70b" else heads,
                          num_hidden_layers=L, vocab_size=32000 if H >= 768 else 512, max_position_embeddings=2048)
        model = LlamaForCausalLM(cfg)
    return model.half().eval()


def synthetic_statistics(model, net, prefix, seed=1234):
    """Synthetic per-channel activation statistics for the LET initialisation (stand-in for
    generate_act_scale_shift.py, which needs real data)."""
    g = torch.Generator().manual_seed(seed)
    H = model.config.hidden_size
    scales, shifts = {}, {}
    names = ["self_attn.q_proj", "self_attn.o_proj", "mlp.up_proj"] if "llama" in net.lower() else \
            ["self_attn.q_proj", "self_attn.out_proj", "fc1"]
    for i in range(model.config.num_hidden_layers):
        for n in names:
            scales[f"{prefix}.{i}.{n}"] = torch.randn(H, generator=g).abs() * 2 + 1e-5
            shifts[f"{prefix}.{i}.{n}"] = torch.randn(H, generator=g) * 0.05
    return scales, shifts


def main(argv=None):
    args = build_parser().parse_args(argv)
    random.seed(args.seed)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    torch.cuda.manual_seed(args.seed)
    if args.epochs > 0:
        assert args.lwc or args.let            # main.py:238-239
    for d in (args.output_dir, args.cache_dir, args.save_dir):
        if d:
            Path(d).mkdir(parents=True, exist_ok=True)
    logger = utils.create_logger(args.output_dir)
    logger.info(args)
    if args.net is None:
        args.net = args.model.split("/")[-1]
    args.model_family = args.net.split("-")[0]
    quant_param_dicts(args)
    dev = torch.device("cuda")
    if args.synthetic or not os.path.isdir(args.model or ""):
        model = synthetic_model(args.net, args.num_layers)
        prefix = "model.layers" if "llama" in args.net.lower() else "model.decoder.layers"
        act_scales, act_shifts = synthetic_statistics(model, args.net, prefix)
        if args.online_act_stats:
            act_scales = act_shifts = None
        g = torch.Generator().manual_seed(1234)
        dataloader = [(torch.randint(0, model.config.vocab_size, (1, args.seqlen), generator=g), None)
                      for _ in range(args.nsamples)]
    else:
        from transformers import AutoModelForCausalLM
        model = AutoModelForCausalLM.from_pretrained(args.model, torch_dtype=torch.float16).eval()
        sc_path, sh_path = args.act_scales or f"./act_scales/{args.net}.pt", args.act_shifts or f"./act_shifts/{args.net}.pt"
        act_scales = act_shifts = None
        if args.let and not args.online_act_stats and os.path.exists(sc_path) and os.path.exists(sh_path):
            act_scales, act_shifts = torch.load(sc_path), torch.load(sh_path)
        elif args.let:
            logger.info("LET activation statistics: collected online in the FP pass of each block")
        cache = f"{args.cache_dir}/dataloader_{args.model_family}_{args.calib_dataset}_{args.nsamples}.cache"
        if not os.path.exists(cache):
            raise FileNotFoundError(f"{cache}: calibration tokens must be prepared offline (datautils is out of scope)")
        dataloader = torch.load(cache)
    for p in model.parameters():
        p.requires_grad = False
    lm = SimpleNamespace(model=model, device=dev, _device=dev, seqlen=args.seqlen)
    if args.wbits < 16 or args.abits < 16:
        logger.info("=== start quantization ===")
        tick = time.time()
        omniquant(lm, args, dataloader, act_scales, act_shifts, logger)
        logger.info(time.time() - tick)
    if args.eval_ppl:
        # main.py:97-148 on a caller-supplied / synthetic token stream (dataset loaders are out of scope)
        from .evaluate import evaluate_ppl
        if args.synthetic or not os.path.isdir(args.model or ""):
            g = torch.Generator().manual_seed(4321)
            testenc = torch.randint(0, model.config.vocab_size, (1, args.seqlen * max(1, args.limit if args.limit > 0 else 4)), generator=g)
        else:
            path = f"{args.cache_dir}/testloader_{args.model_family}_wikitext2_all.cache"
            if not os.path.exists(path):
                raise FileNotFoundError(f"{path}: evaluation tokens must be prepared offline (datautils is out of scope)")
            testenc = torch.load(path)
            testenc = testenc.input_ids if hasattr(testenc, "input_ids") else testenc
        args.ppl = evaluate_ppl(lm, args, testenc, logger, limit=args.limit)
    if args.save_dir:
        model.save_pretrained(args.save_dir)
    return model


if __name__ == "__main__":
    main()
