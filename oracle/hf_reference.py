"""The reference's arithmetic, assembled from the REAL transformers modules (test infrastructure only: imported by tests/,
__graft_entry__.smoke() and bench.py's CPU / HF baseline legs, never by the product package).

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


class _construct_on:
    """Build an HF module directly on `device` without running its random initialiser (the weights are loaded right after).
    At clip-flant5-xxl / Qwen2.5-VL-7B width the default CPU construction + init of 8-11 B parameters takes minutes."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.ctx = []

    _INIT_FNS = ("kaiming_uniform_", "kaiming_normal_", "uniform_", "normal_", "trunc_normal_", "xavier_uniform_", "xavier_normal_")

    def __enter__(self):
        self.ctx = [torch.device(self.device)]
        try:
            from transformers import initialization as hf_init
            self.ctx.append(hf_init.no_init_weights())
        except Exception:   # older / newer transformers: fall back to the (device-side, still fast) default initialiser
            pass
        for c in self.ctx:
            c.__enter__()
        # nn.Linear / nn.Embedding constructors run their own reset_parameters(): on the host that is a single-threaded random fill of
        # every weight. Make the fills no-ops while constructing (the tensors stay untouched virtual memory until the real weights land).
        import torch.nn.init as tinit
        self.saved = {n: getattr(tinit, n) for n in self._INIT_FNS if hasattr(tinit, n)}
        for n in self.saved:
            setattr(tinit, n, lambda t, *a, **k: t)
        return self

    def __exit__(self, *a):
        import torch.nn.init as tinit
        for n, f in self.saved.items():
            setattr(tinit, n, f)
        for c in reversed(self.ctx):
            c.__exit__(*a)
        return False


def _cast_keep_float_buffers(module, dtype, device):
    """module.to(dtype) would also round float BUFFERS (rotary inv_freq) that `from_pretrained(torch_dtype=bf16)` leaves in fp32."""
    keep = {n: b.detach().clone() for n, b in module.named_buffers() if b.is_floating_point()}
    module.to(device=device, dtype=dtype)
    for n, b in keep.items():
        mod = module
        *path, leaf = n.split(".")
        for part in path:
            mod = getattr(mod, part)
        mod._buffers[leaf] = b.to(device)
    return module


def build_hf_modules(cfg, sd: Dict[str, torch.Tensor], dtype=torch.float32, device="cpu", fast_construct=None, assign=False):
    """cfg: oracle.clipt5_oracle.ClipT5Config (or the engine's, same fields). Returns (vision, projector, t5)."""
    from transformers import CLIPVisionConfig, CLIPVisionModel, T5Config, T5ForConditionalGeneration
    big = torch.device(device).type == "cuda" if fast_construct is None else fast_construct
    if big:
        # construct every module on the GPU in the target dtype, uninitialised: load_state_dict below fills everything that is used
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            with _construct_on(device):
                return _build_hf_modules(cfg, sd, dtype, device, CLIPVisionConfig, CLIPVisionModel, T5Config, T5ForConditionalGeneration, assign)
        finally:
            torch.set_default_dtype(prev)
    return _build_hf_modules(cfg, sd, dtype, device, CLIPVisionConfig, CLIPVisionModel, T5Config, T5ForConditionalGeneration, assign)


def _build_hf_modules(cfg, sd, dtype, device, CLIPVisionConfig, CLIPVisionModel, T5Config, T5ForConditionalGeneration, assign=False):
    # assign=True: the modules adopt the tensors of `sd` (already in `dtype` on `device`) instead of copying them: one copy of an
    # 11 B-parameter model in memory

    vcfg = CLIPVisionConfig(hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_mlp, num_hidden_layers=cfg.vit_layers,
                            num_attention_heads=cfg.vit_heads, image_size=cfg.image_size, patch_size=cfg.patch_size,
                            hidden_act="quick_gelu", layer_norm_eps=cfg.vit_ln_eps, attn_implementation="eager")
    vision = CLIPVisionModel(vcfg)
    vsd = {k[len("vision_tower."):]: v for k, v in sd.items() if k.startswith("vision_tower.")}
    missing, unexpected = vision.load_state_dict(vsd, strict=False, assign=assign)
    missing = [m for m in missing if "post_layernorm" not in m and "position_ids" not in m]
    assert not missing and not unexpected, (missing, unexpected)

    proj = nn.Sequential(nn.Linear(cfg.vit_hidden, cfg.d_model), nn.GELU(), nn.Linear(cfg.d_model, cfg.d_model))
    proj.load_state_dict({k[len("mm_projector."):]: v for k, v in sd.items() if k.startswith("mm_projector.")}, assign=assign)

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
    missing, unexpected = t5.load_state_dict(tsd, strict=False, assign=assign)
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
                    feats=feats.float().cpu(), proj=img.float().cpu(), embeds=embeds.float().cpu(), mask=mask.bool().cpu(),
                    dec_hidden=out.decoder_hidden_states[-1].float(),      # after the final T5LayerNorm = lm_head input (stays on the device)
                    logprobs=torch.log_softmax(logits.float(), -1).gather(-1, lab.clamp(min=0)[..., None])[..., 0].cpu())
    return scores


# ================================================================================================ Qwen2.5-VL (reference path)
def hf_qwen_config(cfg):
    """oracle.qwen25vl_oracle.Qwen25VLConfig -> transformers Qwen2_5_VLConfig (eager attention, untied lm_head)."""
    from transformers import Qwen2_5_VLConfig
    c = Qwen2_5_VLConfig(
        text_config=dict(vocab_size=cfg.vocab, hidden_size=cfg.hidden, intermediate_size=cfg.mlp, num_hidden_layers=cfg.layers,
                         num_attention_heads=cfg.heads, num_key_value_heads=cfg.kv_heads, rms_norm_eps=cfg.rms_eps,
                         rope_parameters=dict(rope_type="default", rope_theta=cfg.rope_theta, mrope_section=list(cfg.mrope_section)),
                         tie_word_embeddings=False, max_position_embeddings=4096, use_sliding_window=False),
        vision_config=dict(depth=cfg.vit_depth, hidden_size=cfg.vit_hidden, intermediate_size=cfg.vit_mlp, num_heads=cfg.vit_heads,
                           patch_size=cfg.patch_size, temporal_patch_size=cfg.temporal_patch_size,
                           spatial_merge_size=cfg.spatial_merge_size, window_size=cfg.window_size,
                           fullatt_block_indexes=list(cfg.fullatt_block_indexes), out_hidden_size=cfg.out_hidden,
                           tokens_per_second=cfg.tokens_per_second, hidden_act="silu"),
        image_token_id=cfg.image_token_id, video_token_id=cfg.video_token_id, tie_word_embeddings=False)
    c._attn_implementation = "eager"
    return c


def build_hf_qwen(cfg, sd: Dict[str, torch.Tensor], dtype=torch.float32, device="cpu", attn="eager", fast_construct=None, assign=False):
    """The real Qwen2_5_VLForConditionalGeneration carrying `sd` (HF in-memory names). On a GPU it is constructed there directly,
    uninitialised, in `dtype` (what from_pretrained(torch_dtype=bf16) does), so rotary inv_freq buffers stay fp32."""
    from transformers import Qwen2_5_VLForConditionalGeneration
    c = hf_qwen_config(cfg)
    c._attn_implementation = attn
    big = torch.device(device).type == "cuda" if fast_construct is None else fast_construct
    if big:
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            with _construct_on(device):
                m = Qwen2_5_VLForConditionalGeneration(c)
        finally:
            torch.set_default_dtype(prev)
        missing, unexpected = m.load_state_dict(sd, strict=False, assign=assign)   # assign: adopt `sd`'s tensors (one copy in memory)
    else:
        m = Qwen2_5_VLForConditionalGeneration(c)
        missing, unexpected = m.load_state_dict({k: v.float() for k, v in sd.items()}, strict=False)
        _cast_keep_float_buffers(m, dtype, device)
    assert not missing and not unexpected, (missing, unexpected)
    return m.eval().requires_grad_(False)


@torch.no_grad()
def hf_qwen_reference_scores(model, cfg, pixel_patches, grid_thw, input_ids, answer_ids, temperature: float = 1.0,
                             video=False, second_per_grid_ts=None, return_hidden=False):
    """The reference's scoring recipe, sample by sample (t2v_metrics/models/vqascore_models/qwen2vl_model.py:190-289):
    generate(max_new_tokens=1, temperature=1.0, do_sample=False, output_scores=True, return_dict_in_generate=True), then
    softmax(scores[-1][0] / temperature)[answer_token] (:160-167). Inputs are what the HF processor would hand over: the sample's
    patch rows, its grid, ids with the image-token run, attention mask of ones, mm_token_type_ids."""
    dev = next(model.parameters()).device
    wdtype = next(model.parameters()).dtype
    probs, hidden, offs = [], [], [0]
    for t, h, w in grid_thw:
        offs.append(offs[-1] + t * h * w)
    for b, ids in enumerate(input_ids):
        ids = torch.as_tensor(ids, dtype=torch.long, device=dev)
        px = pixel_patches[offs[b]:offs[b + 1]].to(dev, wdtype)
        grid = torch.tensor([list(grid_thw[b])], device=dev)
        kw = dict(input_ids=ids[None], attention_mask=torch.ones(1, len(ids), dtype=torch.long, device=dev))
        if video:
            kw.update(pixel_values_videos=px, video_grid_thw=grid, mm_token_type_ids=torch.where(ids == cfg.video_token_id, 2, 0)[None])
            if second_per_grid_ts is not None:
                kw["second_per_grid_ts"] = torch.tensor([second_per_grid_ts[b]], device=dev)
        else:
            kw.update(pixel_values=px, image_grid_thw=grid, mm_token_type_ids=(ids == cfg.image_token_id).long()[None])
        out = model.generate(**kw, max_new_tokens=1, temperature=1.0, do_sample=False, output_scores=True,
                             return_dict_in_generate=True, output_hidden_states=return_hidden)
        logits = out.scores[-1][0]
        probs.append(torch.softmax(logits / temperature, dim=-1)[int(answer_ids[b])].float().cpu())
        if return_hidden:
            hidden.append(out.hidden_states[0][-1][0, -1].float())     # prefill step, last layer (after the final norm), last position
    p = torch.stack(probs)
    return (p, torch.stack(hidden)) if return_hidden else p


@torch.no_grad()
def calibrate_rows(hidden: torch.Tensor, lm_head: torch.Tensor, token_id: int, offsets: torch.Tensor) -> torch.Tensor:
    """Fixture helper shared by the full-width tests: pick lm_head[token_id] (least norm, bf16-representable) so that for final hidden
    states hidden[b] the answer logit is LSE(other logits) + offsets[b], i.e. P(answer) = sigmoid(offsets[b]): scores spread over
    (0.1, 0.9) instead of sitting at 1/vocab, so that an absolute 1e-3 tolerance on the probability means something."""
    H = hidden.float()
    others = H @ lm_head.float().t()
    others[:, token_id] = float("-inf")
    target = torch.logsumexp(others, dim=-1) + offsets.to(H.device)
    w = torch.linalg.pinv(H.double()) @ target.double()
    return w.float().to(torch.bfloat16)
