"""Block-wise calibration driver with the reference's entry point
``omniquant(lm, args, dataloader, act_scales, act_shifts, logger)`` (reference
quantize/omniquant.py:22-289) on the B200 block engine.

Per block: FP targets -> LET init -> AdamW loop -> in-place fold + quantise -> re-quantised
forward -> checkpoint, with the same loop semantics, hyper-parameters, parameter / checkpoint
key names (``omni_parameters.pth``: {block_idx: OrderedDict(name -> tensor)}).  Differences by
design: bf16 storage instead of fp16 autocast (no GradScaler), losses stay on the
device (no per-step ``.item()``), the step can be replayed as a CUDA graph, and when
``torch.distributed`` is initialised the calibration samples are sharded over the ranks with one
all-reduce of the flat LWC/LET gradient buffer per step (== the reference with
``--batch_size world_size``).  ``--real_quant`` packs the calibrated block into the AutoGPTQ wire format on the GPU
(quantize/real_quant.py, csrc/gptq_pack.cu; reference :255-278).
"""
import os

import torch
import torch.nn as nn

from .. import engine as E
from ..parallel import make_allreduce, shard_indices, world_info
from ..models.int_llama_layer import QuantLlamaDecoderLayer
from ..models.int_opt_layer import QuantOPTDecoderLayer
from ..models.int_falcon_layer import QuantFalconDecoderLayer
from .int_linear import QuantLinear


def get_named_linears(module):
    return {name: m for name, m in module.named_modules() if isinstance(m, QuantLinear)}


class _StopForward(Exception):
    pass


def _family(model, net):
    net = net.lower()
    if "llama" in net:
        return dict(name="llama", layers=model.model.layers, cls=QuantLlamaDecoderLayer,
                    embed=[model.model.embed_tokens, model.model.norm], prefix="model.layers",
                    pairs={"q_proj": "qkv", "o_proj": "out", "up_proj": "fc1"})
    if "opt" in net:
        dec = model.model.decoder
        embed = [dec.embed_tokens, dec.embed_positions]
        for extra in ("project_out", "project_in"):
            if getattr(dec, extra, None):
                embed.append(getattr(dec, extra))
        return dict(name="opt", layers=dec.layers, cls=QuantOPTDecoderLayer, embed=embed, prefix="model.decoder.layers",
                    pairs={"q_proj": "qkv", "out_proj": "out", "fc1": "fc1"})
    if "falcon" in net:
        return dict(name="falcon", layers=model.transformer.h, cls=QuantFalconDecoderLayer,
                    embed=[model.transformer.word_embeddings, model.transformer.ln_f], prefix="model.transformer.h", pairs={})
    raise ValueError("Only support for opt/llama/Llama-2/falcon now")


def capture_block_inputs(model, layers, dataloader, nsamples, seqlen, dev, dtype):
    """Run the embeddings and intercept the first block's input (the reference's Catcher, :89-113)."""
    H = model.config.hidden_size
    inps = torch.zeros((nsamples, seqlen, H), dtype=dtype, device=dev)
    seen = {"i": 0, "attention_mask": None, "position_ids": None}

    class Catcher(nn.Module):
        def __init__(self, module):
            super().__init__()
            self.module = module

        def forward(self, inp, *a, **kwargs):
            inps[seen["i"]] = inp.to(dtype)
            seen["i"] += 1
            seen["attention_mask"] = kwargs.get("attention_mask")
            seen["position_ids"] = kwargs.get("position_ids")
            raise _StopForward

    layers[0] = Catcher(layers[0])
    with torch.no_grad():
        for batch in dataloader:
            if seen["i"] >= nsamples:
                break
            try:
                model(batch[0].to(dev))
            except _StopForward:
                pass
    layers[0] = layers[0].module
    return inps, seen["attention_mask"], seen["position_ids"]


def init_let_parameters(qlayer, layer_idx, fam, args, act_scales, act_shifts, dev):
    """Channel-wise scale / shift initialisation (:182-197): scale = act^alpha / max_k(W)^(1-alpha), shift from the
    activation statistics (zero for LLaMA), qkt scale = 1."""
    is_llama = fam["name"] == "llama"
    nq = get_named_linears(qlayer)
    q_out = next(m for n, m in nq.items() if n.endswith("q_proj")).out_features
    qlayer.register_parameter("qkt_smooth_scale", nn.Parameter(torch.ones(q_out, device=dev, dtype=torch.float32)))
    for name, module in nq.items():
        for key, tag in fam["pairs"].items():
            if key in name:
                full = f"{fam['prefix']}.{layer_idx}.{name}"
                # the reference evaluates this in the model's dtype (fp16 under its default autocast path, :189-191; fp32
                # with --deactive_amp) before the parameter is promoted by qlayer.float(); same rounding here
                idt = torch.float32 if getattr(args, "deactive_amp", False) else torch.float16
                act = act_scales[full].to(device=dev, dtype=idt).clamp(min=1e-5)
                weight = module.weight.to(idt).max(dim=0)[0].clamp(min=1e-5)       # max, not abs().max() (:190)
                scale = (act.pow(args.alpha) / weight.pow(1 - args.alpha)).clamp(min=1e-5).float()
                if is_llama:
                    shift = torch.zeros_like(scale)
                else:
                    shift = act_shifts[full].to(device=dev, dtype=torch.float32)
                qlayer.register_parameter(f"{tag}_smooth_shift", nn.Parameter(shift))
                qlayer.register_parameter(f"{tag}_smooth_scale", nn.Parameter(scale))


def _stat_names(qlayer, layer_idx, fam):
    """{reference key 'prefix.i.linear_name': LET tag} of the linears whose input statistics initialise LET."""
    out = {}
    for name in get_named_linears(qlayer):
        for key, tag in fam["pairs"].items():
            if key in name:
                out[f"{fam['prefix']}.{layer_idx}.{name}"] = tag
    return out


def omniquant(lm, args, dataloader, act_scales, act_shifts, logger=None, inps=None):
    """Calibrate every decoder block of ``lm.model``.  ``inps`` (optional, [nsamples, seqlen, H]) bypasses the
    embedding capture for synthetic runs."""
    log = logger.info if logger is not None else (lambda *a, **k: None)
    log("Starting ...")
    model, dev = lm.model, lm.device
    use_cache = model.config.use_cache
    model.config.use_cache = False
    fam = _family(model, args.net)
    layers = fam["layers"]
    rank, world = world_info()
    if args.batch_size % world != 0:
        raise ValueError(f"--batch_size {args.batch_size} must be a multiple of the world size {world}")
    b_local = args.batch_size // world
    if inps is None:
        for m in fam["embed"]:
            m.to(dev)
        layers[0] = layers[0].to(dev)
        inps, _, _ = capture_block_inputs(model, layers, dataloader, args.nsamples, lm.seqlen, dev, torch.bfloat16)
        layers[0] = layers[0].cpu()
        for m in fam["embed"]:
            m.cpu()
        torch.cuda.empty_cache()
    inps = inps.to(dev, torch.bfloat16)
    # sample sharding: global step j consumes samples [j*B, (j+1)*B); rank r owns b_local of them
    idx = shard_indices(args.nsamples, args.batch_size, rank, world)
    quant_inps = inps[idx].contiguous()
    fp_inps = quant_inps.clone()
    del inps
    allreduce = make_allreduce()
    omni_parameters = torch.load(args.resume) if getattr(args, "resume", None) else {}
    T = lm.seqlen
    use_graph = getattr(args, "cuda_graph", True)

    from ..streaming import BlockStreamer
    streamer = BlockStreamer(layers, dev, enabled=getattr(args, "prefetch_blocks", True))

    for i in range(len(layers)):
        log(f"=== Start quantize layer {i} ===")
        layer = streamer.fetch(i)                  # :159, already resident when block i-1 prefetched it
        streamer.prefetch(i + 1)                   # next block's H2D copy overlaps this block's calibration
        qlayer = fam["cls"](model.config, layer, args).to(dev)
        qlayer.set_quant_state(weight_quant=False, act_quant=True)
        qlayer.let = args.let
        online = args.let and act_scales is None      # no statistics file: collect them in the FP pass of this block
        if args.let:
            if online:
                H = model.config.hidden_size
                names = _stat_names(qlayer, i, fam)
                init_let_parameters(qlayer, i, fam, args, {k: torch.ones(H) for k in names},
                                    {k: torch.zeros(H) for k in names}, dev)      # placeholders until the FP pass ran
            else:
                init_let_parameters(qlayer, i, fam, args, act_scales, act_shifts, dev)
        if getattr(args, "resume", None):
            qlayer.load_state_dict(omni_parameters[i], strict=False)
        qlayer.float()
        eng = qlayer._ensure_engine(T, b_local)
        eng.world = world
        if args.epochs > 0:
            collect = online and not getattr(args, "resume", None)
            stats = E.new_act_stats(eng, fp_inps.shape[0]) if collect else None
            fp2 = E.fp_pass(eng, fp_inps, quant_inps, args.aug_loss, stats=stats)
            if collect:
                folded = E.fold_act_stats(stats, idx, args.nsamples, allreduce)
                names = _stat_names(qlayer, i, fam)
                init_let_parameters(qlayer, i, fam, args, {k: folded[t][0] for k, t in names.items()},
                                    {k: folded[t][1] for k, t in names.items()}, dev)
                qlayer.float()
                qlayer._push_params()
                if getattr(args, "save_act_stats", None) is not None:
                    args.save_act_stats[0].update({k: folded[t][0].cpu() for k, t in names.items()})
                    args.save_act_stats[1].update({k: folded[t][1].cpu() for k, t in names.items()})

            def epoch_log(ep, ls, i=i):
                # the per-step losses were all-reduced with the gradients, so every rank sees the same values and all
                # ranks abort together (the reference checks `loss.isfinite()` per step, :223-225)
                ls = ls / world
                if not bool(torch.isfinite(ls).all()):
                    raise FloatingPointError(f"layer {i} epoch {ep}: loss is not finite")
                if logger is not None:
                    log(f"layer {i} iter {ep} loss:{ls.mean().item()} max memory_allocated "
                        f"{torch.cuda.max_memory_allocated(dev) / 1024 ** 2} ")

            E.train(eng, fp_inps, quant_inps, epochs=args.epochs, fp_inps_2=fp2, allreduce=allreduce, log=epoch_log,
                    use_graph=use_graph)
            qlayer._pull_params()
            qlayer.clear_temp_variable()
        qlayer.smooth_and_quant_inplace()          # real smooth and quantization (:239)
        if args.epochs > 0:
            E.requant_pass(eng, quant_inps)        # update input of quantization model (:242-245)
        qlayer.register_scales_and_zeros()
        if os.environ.get("OQ_DEBUG_STREAMS") == "1":
            f = lambda t: f"absmax {float(t.float().abs().max()):.3e} rms {float(t.float().pow(2).mean().sqrt()):.3e} finite {bool(torch.isfinite(t).all())}"
            print(f"[oq debug] after layer {i}: fp stream {f(fp_inps)} | quant stream {f(quant_inps)}", flush=True)
        qlayer._eng = None                         # drop the fp32 masters / workspaces of this block
        qlayer._final_ws = None
        saved = None
        if getattr(args, "real_quant", False):     # pack before the fp16 cast: same integer codes, on the GPU (:255-278)
            from .real_quant import pack_block
            saved = {k: v.detach().half().cpu() for k, v in qlayer.omni_state_dict().items()}   # LWC factors leave with the QuantLinears
            pack_block(qlayer, args.wbits)
        qlayer.half()
        layers[i] = streamer.writeback(qlayer)     # :248 `qlayer.to("cpu")`, asynchronous: overlaps the next block
        if args.epochs > 0:
            omni_parameters[i] = saved if saved is not None else qlayer.omni_state_dict()
            if rank == 0 and getattr(args, "output_dir", None):
                streamer.drain()               # the checkpoint reads the host copies
                torch.save(omni_parameters, os.path.join(args.output_dir, "omni_parameters.pth"))
        del layer, eng

    del quant_inps, fp_inps
    streamer.drain()               # the host copies of the calibrated blocks are complete from here on
    from ..models._block import release_engines
    release_engines()              # the resident engine + its step graph (holds the NCCL all-reduce under DP)
    torch.cuda.synchronize()
    model.config.use_cache = use_cache
    return model
