"""Post-calibration fake-quant inference: perplexity of the calibrated model on a token stream.

Reference: main.py:97-148 (the wikitext2 / ptb / c4 perplexity loop of ``evaluate``): non-overlapping ``seqlen``
windows, model body -> final norm -> lm_head, shifted cross entropy, ``ppl = exp(sum(nll) / (nsamples * seqlen))``.
The decoder blocks are the calibrated Quant*DecoderLayer modules (weights already fake-quantised in place, dynamic
per-token activation quantisation still on, ``set_quant_state(False, True)`` as quantize/omniquant.py:239-246 leaves
them); they run on the same CUDA block engine as the calibration, one block resident at a time like the reference's
``--multigpu``-less path for big models (main.py:65-95 moves the pieces it needs).  Dataset loading (datautils) and
lm_eval tasks are out of scope: the caller supplies the token tensor.
"""
from typing import Optional

import torch
from torch import nn

from .quantize.omniquant import _family, capture_block_inputs


def _head(model, fam_name):
    """(final norm or None, output projection or None, lm_head) of the architecture."""
    if fam_name == "llama":
        return model.model.norm, None, model.lm_head
    if fam_name == "opt":
        dec = model.model.decoder
        return getattr(dec, "final_layer_norm", None), getattr(dec, "project_out", None), model.lm_head
    return model.transformer.ln_f, None, model.lm_head


@torch.no_grad()
def block_chain(lm, args, tokens: torch.Tensor, keep_on_device: bool = False) -> torch.Tensor:
    """Hidden states [n, seqlen, H] (bf16, on the device) after the last decoder block for ``tokens`` [n, seqlen]."""
    model, dev, T = lm.model, lm.device, lm.seqlen
    fam = _family(model, args.net)
    layers = fam["layers"]
    n = tokens.shape[0]
    for m in fam["embed"]:
        m.to(dev)
    hs, mask, _ = capture_block_inputs(model, layers, [(tokens[i:i + 1], None) for i in range(n)], n, T, dev, torch.bfloat16)
    for m in fam["embed"]:
        m.cpu()
    if not (isinstance(mask, torch.Tensor) and mask.dim() == 4 and mask.is_floating_point()):
        mask = None                     # causal by default (the engine builds the mask from seq_len)
    for j in range(len(layers)):
        layer = layers[j]
        if not hasattr(layer, "_run"):
            raise TypeError(f"layer {j} is not a calibrated Quant*DecoderLayer (run omniquant() first)")
        layer = layer.to(dev)
        for i in range(n):
            hs[i:i + 1] = layer(hs[i:i + 1], attention_mask=mask)[0]
        if not keep_on_device:
            layer._eng = None           # drop the engine's device buffers with the block
            layers[j] = layer.cpu()
    return hs


@torch.no_grad()
def evaluate_ppl(lm, args, testenc: torch.Tensor, logger=None, limit: int = -1) -> float:
    """Perplexity over ``testenc`` ([1, n_tokens] or [n, seqlen] token ids), main.py:111-148."""
    model, dev, T = lm.model, lm.device, lm.seqlen
    if testenc.dim() == 2 and testenc.shape[0] == 1:
        n = testenc.numel() // T
        testenc = testenc[0, :n * T].view(n, T)
    if limit > 0:
        testenc = testenc[:limit]
    n = testenc.shape[0]
    if n == 0:
        raise ValueError("evaluate_ppl: fewer tokens than one seqlen window")
    use_cache = model.config.use_cache
    model.config.use_cache = False
    fam = _family(model, args.net)
    hs = block_chain(lm, args, testenc)
    norm, proj, head = _head(model, fam["name"])
    mods = [m for m in (norm, proj, head) if m is not None]
    for m in mods:
        m.to(dev)
    loss_fct = nn.CrossEntropyLoss()
    nlls = []
    for i in range(n):
        h = hs[i:i + 1].to(head.weight.dtype)
        if norm is not None:
            h = norm(h)
        if proj is not None:
            h = proj(h)
        logits = head(h)
        shift_logits = logits[:, :-1, :].float()
        shift_labels = testenc[i:i + 1, 1:].to(dev)
        loss = loss_fct(shift_logits.view(-1, shift_logits.size(-1)), shift_labels.reshape(-1))
        nlls.append(loss.float() * T)                      # main.py:141-142
    for m in mods:
        m.cpu()
    ppl = torch.exp(torch.stack(nlls).sum() / (n * T))     # main.py:146
    model.config.use_cache = use_cache
    if logger is not None:
        logger.info(f"perplexity : {ppl.item()}")
    return float(ppl.item())
