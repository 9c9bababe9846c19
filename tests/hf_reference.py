"""The reference's arithmetic, assembled from the REAL transformers modules (test infrastructure only).

The v3.0 CLIP-FlanT5 wrapper is gone from the reference snapshot, but the arithmetic it delegates to is
`transformers` (T5ForConditionalGeneration + CLIPVisionModel). This module composes those classes exactly as the
wrapper did: vision tower hidden_states[-2][:,1:] -> mlp2x_gelu projector -> splice at the -200 slot ->
T5 forward with labels -> exp(-CE). It is used to (a) pin oracle/clipt5_oracle.py, (b) generate tests/golden/*.pt
(tools/make_golden.py), (c) run the reference forward on the GPU box next to the engine (tests -m gpu).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

IMAGE_TOKEN_INDEX = -200
IGNORE_INDEX = -100


def build_hf_modules(cfg, sd: Dict[str, torch.Tensor], dtype=torch.float32, device="cpu"):
    """cfg: oracle.clipt5_oracle.ClipT5Config (or the engine's, same fields). Returns (vision, projector, t5)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel, T5Config, T5ForConditionalGeneration

    vcfg = CLIPVisionConfig(hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_mlp, num_hidden_layers=cfg.vit_layers,
                            num_attention_heads=cfg.vit_heads, image_size=cfg.image_size, patch_size=cfg.patch_size,
                            hidden_act="quick_gelu", layer_norm_eps=cfg.vit_ln_eps, attn_implementation="eager")
    vision = CLIPVisionModel(vcfg)
    vsd = {k[len("vision_tower."):]: v for k, v in sd.items() if k.startswith("vision_tower.")}
    missing, unexpected = vision.load_state_dict(vsd, strict=False)
    missing = [m for m in missing if "post_layernorm" not in m and "position_ids" not in m]
    assert not missing and not unexpected, (missing, unexpected)

    proj = nn.Sequential(nn.Linear(cfg.vit_hidden, cfg.d_model), nn.GELU(), nn.Linear(cfg.d_model, cfg.d_model))
    proj.load_state_dict({k[len("mm_projector."):]: v for k, v in sd.items() if k.startswith("mm_projector.")})

    tcfg = T5Config(vocab_size=cfg.vocab, d_model=cfg.d_model, d_kv=cfg.d_kv, d_ff=cfg.d_ff, num_layers=cfg.enc_layers,
                    num_decoder_layers=cfg.dec_layers, num_heads=cfg.n_heads,
                    relative_attention_num_buckets=cfg.rel_buckets, relative_attention_max_distance=cfg.rel_max_distance,
                    dropout_rate=0.0, layer_norm_epsilon=cfg.t5_ln_eps, feed_forward_proj="gated-gelu",
                    tie_word_embeddings=False, pad_token_id=cfg.pad_token_id, eos_token_id=1,
                    decoder_start_token_id=cfg.decoder_start_id, use_cache=False)
    t5 = T5ForConditionalGeneration(tcfg)
    # transformers 5.5 force-ties lm_head to `shared` even with tie_word_embeddings=False: untie explicitly.
    t5.lm_head = nn.Linear(cfg.d_model, cfg.vocab, bias=False)
    tsd = {k: v for k, v in sd.items() if not k.startswith(("vision_tower.", "mm_projector."))}
    tsd["encoder.embed_tokens.weight"] = sd["shared.weight"]
    tsd["decoder.embed_tokens.weight"] = sd["shared.weight"]
    missing, unexpected = t5.load_state_dict(tsd, strict=False)
    assert not unexpected and not [m for m in missing if "embed_tokens" not in m], (missing, unexpected)
    assert t5.lm_head.weight.data_ptr() != t5.shared.weight.data_ptr()
    assert getattr(t5.config, "scale_decoder_outputs", False) is False
    for m in (vision, proj, t5):
        m.to(device=device, dtype=dtype).eval().requires_grad_(False)
    return vision, proj, t5


@torch.no_grad()
def hf_clipt5_forward(cfg, modules, pixels, input_ids, text_lens, labels, image_index: Optional[torch.Tensor] = None,
                      autocast_bf16: bool = False, return_all: bool = False):
    """v3.0 CLIPT5Model.forward restated around the HF modules."""
    vision, proj, t5 = modules
    dev = next(t5.parameters()).device
    wdtype = next(t5.parameters()).dtype
    pixels = pixels.to(dev)
    ctx = torch.autocast(device_type=dev.type, dtype=torch.bfloat16, enabled=autocast_bf16)
    with ctx:
        vout = vision(pixels.to(wdtype), output_hidden_states=True)
        feats = vout.hidden_states[-2][:, 1:]
        img = proj(feats.to(wdtype))
        B, L = input_ids.shape
        P = img.shape[1]
        S = L - 1 + P
        embeds = torch.zeros(B, S, cfg.d_model, dtype=img.dtype, device=dev)
        mask = torch.zeros(B, S, dtype=torch.long, device=dev)
        for b in range(B):
            ids = input_ids[b, : int(text_lens[b])].to(dev)
            pos = (ids == IMAGE_TOKEN_INDEX).nonzero()
            im = img[int(image_index[b]) if image_index is not None else b]
            if len(pos) == 0:
                seq = t5.shared(ids)
            else:
                s = int(pos[0])
                seq = torch.cat([t5.shared(ids[:s]).to(img.dtype), im, t5.shared(ids[s + 1:]).to(img.dtype)], dim=0)
            embeds[b, : seq.shape[0]] = seq
            mask[b, : seq.shape[0]] = 1
        lab = labels.to(dev)
        out = t5(inputs_embeds=embeds, attention_mask=mask, labels=lab, decoder_attention_mask=(lab != IGNORE_INDEX).long(),
                 output_hidden_states=return_all)
        logits = out.logits
        scores = torch.zeros(B)
        loss_fct = nn.CrossEntropyLoss(reduction="mean")
        for k in range(B):
            scores[k] = (-loss_fct(logits[k], lab[k])).exp()
    if return_all:
        return dict(scores=scores, logits=logits.float().cpu(), enc=out.encoder_last_hidden_state.float().cpu(),
                    feats=feats.float().cpu(), proj=img.float().cpu(), embeds=embeds.float().cpu(), mask=mask.bool().cpu())
    return scores
