"""Seeded synthetic weights / inputs generated directly on the GPU in the engine's fused layout.

There is no network and no checkpoint on the build or GPU boxes, so bench.py and the full-size tests run the real
architecture (clip-flant5-xxl dims) with random-init weights, as BASELINE.json prescribes ("data": "synthetic").
Scales follow the HF initialisers (T5 `_init_weights`, modeling_t5.py:541-593, factor 1.0; CLIP std 0.02) so activations
stay O(1) through 24+24 layers; lm_head ~ N(0, 1/d_model) keeps logits O(1).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .config import ClipT5Config


def synthetic_engine_weights(cfg: ClipT5Config, device, seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}

    def nrm(name, *shape, std=1.0):
        out[name] = (torch.randn(*shape, generator=g, device=device, dtype=torch.bfloat16) * std).contiguous()

    def gain(name, n):
        out[name] = (1.0 + 0.1 * torch.randn(n, generator=g, device=device)).to(torch.bfloat16)

    Dv, Dm, dff, inner = cfg.vit_hidden, cfg.d_model, cfg.d_ff, cfg.n_heads * cfg.d_kv
    k_real = 3 * cfg.patch_size ** 2
    kpad = (k_real + 63) // 64 * 64
    pw = torch.zeros(Dv, kpad, device=device)
    pw[:, :k_real] = torch.randn(Dv, k_real, generator=g, device=device) * 0.02
    out["vit.patch_embed.weight"] = pw.to(torch.bfloat16)
    nrm("vit.class_embedding", Dv, std=Dv ** -0.5)
    nrm("vit.position_embedding", cfg.num_patches + 1, Dv, std=0.02)
    gain("vit.pre_ln.weight", Dv); nrm("vit.pre_ln.bias", Dv, std=0.02)
    for l in range(cfg.vit_layers - 1):
        p = f"vit.layers.{l}."
        gain(p + "ln1.weight", Dv); nrm(p + "ln1.bias", Dv, std=0.02)
        gain(p + "ln2.weight", Dv); nrm(p + "ln2.bias", Dv, std=0.02)
        nrm(p + "qkv.weight", 3 * Dv, Dv, std=Dv ** -0.5); nrm(p + "qkv.bias", 3 * Dv, std=0.02)
        nrm(p + "out.weight", Dv, Dv, std=Dv ** -0.5); nrm(p + "out.bias", Dv, std=0.02)
        nrm(p + "fc1.weight", cfg.vit_mlp, Dv, std=Dv ** -0.5); nrm(p + "fc1.bias", cfg.vit_mlp, std=0.02)
        nrm(p + "fc2.weight", Dv, cfg.vit_mlp, std=cfg.vit_mlp ** -0.5); nrm(p + "fc2.bias", Dv, std=0.02)
    nrm("proj.0.weight", Dm, Dv, std=Dv ** -0.5); nrm("proj.0.bias", Dm, std=0.02)
    nrm("proj.2.weight", Dm, Dm, std=Dm ** -0.5); nrm("proj.2.bias", Dm, std=0.02)
    nrm("t5.shared", cfg.vocab, Dm, std=1.0)
    nrm("t5.lm_head", cfg.vocab, Dm, std=Dm ** -0.5)
    nrm("t5.enc.rel_bias", cfg.rel_buckets, cfg.n_heads, std=0.5)
    nrm("t5.dec.rel_bias", cfg.rel_buckets, cfg.n_heads, std=0.5)
    gain("t5.enc.final_ln", Dm); gain("t5.dec.final_ln", Dm)

    def qkv(name):
        q = torch.randn(inner, Dm, generator=g, device=device, dtype=torch.bfloat16) * (Dm * cfg.d_kv) ** -0.5
        kv = torch.randn(2 * inner, Dm, generator=g, device=device, dtype=torch.bfloat16) * Dm ** -0.5
        out[name] = torch.cat([q, kv], dim=0).contiguous()

    for l in range(cfg.enc_layers):
        p = f"t5.enc.{l}."
        gain(p + "ln0", Dm); gain(p + "ln1", Dm)
        qkv(p + "qkv")
        nrm(p + "o", Dm, inner, std=inner ** -0.5)
        nrm(p + "wi", 2 * dff, Dm, std=Dm ** -0.5)
        nrm(p + "wo", Dm, dff, std=dff ** -0.5)
    for l in range(cfg.dec_layers):
        p = f"t5.dec.{l}."
        gain(p + "ln0", Dm); gain(p + "ln1", Dm); gain(p + "ln2", Dm)
        qkv(p + "qkv")
        nrm(p + "o", Dm, inner, std=inner ** -0.5)
        nrm(p + "cq", inner, Dm, std=(Dm * cfg.d_kv) ** -0.5)
        nrm(p + "ckv", 2 * inner, Dm, std=Dm ** -0.5)
        out[p + "ckT"] = out[p + "ckv"][:inner].t().contiguous()
        nrm(p + "co", Dm, inner, std=inner ** -0.5)
        nrm(p + "wi", 2 * dff, Dm, std=Dm ** -0.5)
        nrm(p + "wo", Dm, dff, std=dff ** -0.5)
    return out


def synthetic_batch(cfg: ClipT5Config, batch: int, text_len: int = 97, seed: int = 1, ragged: bool = False,
                    label_ids=(2163, 1), source_size: int = 512, n_images: Optional[int] = None, raw_u8: bool = False):
    """HOST tensors of one step of BASELINE config 2: `batch` uniform-random uint8 source_size^2 RGB images, already taken
    through the reference pre-processing's geometry (a square image -> bicubic resize to image_size; for i.i.d. noise the
    resampled pixel statistics, not their values, are what matter to a throughput run, so the bench draws the resized
    pixels directly) and CLIP normalisation; `text_len` ids with one image slot; labels = label_ids. All pinned."""
    g = torch.Generator().manual_seed(seed)
    ni = n_images or batch
    mean = torch.tensor((0.48145466, 0.4578275, 0.40821073))[None, :, None, None]
    std = torch.tensor((0.26862954, 0.26130258, 0.27577711))[None, :, None, None]
    raw = torch.randint(0, 256, (ni, 3, cfg.image_size, cfg.image_size), generator=g, dtype=torch.uint8)
    pixels = ((raw.float() / 255.0) - mean) / std
    ids = torch.randint(2, cfg.vocab - 28, (batch, text_len), generator=g, dtype=torch.int32)
    lens = torch.full((batch,), text_len, dtype=torch.int32)
    if ragged:
        lens = torch.randint(64, text_len + 1, (batch,), generator=g, dtype=torch.int32)
    for b in range(batch):
        n = int(lens[b])
        ids[b, int(torch.randint(1, n - 1, (1,), generator=g))] = -200
        ids[b, n - 1] = 1
        ids[b, n:] = cfg.pad_token_id
    labels = torch.tensor([list(label_ids)] * batch, dtype=torch.int32)
    out = dict(pixels=pixels, input_ids=ids, text_lens=lens, labels=labels)
    if raw_u8:          # the decoded source images themselves (uint8 HWC), for the end-to-end path with device pre-processing
        out["raw_u8"] = torch.randint(0, 256, (ni, source_size, source_size, 3), generator=g, dtype=torch.uint8)
    if torch.cuda.is_available():
        out = {k: v.pin_memory() for k, v in out.items()}
    return out


def synthetic_qwen_engine_weights(cfg, device, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Qwen2.5-VL weights in the engine's fused/padded layout (see engine.convert_qwen_state_dict), generated on the GPU."""
    g = torch.Generator(device=device).manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}

    def nrm(*shape, std=0.02):
        return torch.randn(*shape, generator=g, device=device, dtype=torch.bfloat16) * std

    def gain(n):
        return (1.0 + 0.1 * torch.randn(n, generator=g, device=device)).to(torch.bfloat16)

    Dv, H, hd, mp, mlp_v = cfg.vit_hidden, cfg.vit_heads, cfg.vit_head_dim, cfg.vit_mlp_padded, cfg.vit_mlp
    out["vis.patch_embed"] = nrm(Dv, cfg.patch_dim, std=cfg.patch_dim ** -0.5)
    for l in range(cfg.vit_depth):
        q = f"vis.{l}."
        out[q + "norm1"] = gain(Dv); out[q + "norm2"] = gain(Dv)
        out[q + "qkv.weight"] = nrm(3 * H * hd, Dv, std=Dv ** -0.5); out[q + "qkv.bias"] = nrm(3 * H * hd)
        out[q + "proj.weight"] = nrm(Dv, H * hd, std=Dv ** -0.5); out[q + "proj.bias"] = nrm(Dv)
        gu = torch.zeros(2 * mp, Dv, device=device, dtype=torch.bfloat16)
        gu[:mlp_v] = nrm(mlp_v, Dv, std=Dv ** -0.5); gu[mp:mp + mlp_v] = nrm(mlp_v, Dv, std=Dv ** -0.5)
        gb = torch.zeros(2 * mp, device=device, dtype=torch.bfloat16); gb[:mlp_v] = nrm(mlp_v); gb[mp:mp + mlp_v] = nrm(mlp_v)
        out[q + "gate_up.weight"] = gu; out[q + "gate_up.bias"] = gb
        dw = torch.zeros(Dv, mp, device=device, dtype=torch.bfloat16); dw[:, :mlp_v] = nrm(Dv, mlp_v, std=mlp_v ** -0.5)
        out[q + "down.weight"] = dw; out[q + "down.bias"] = nrm(Dv)
    M = Dv * cfg.spatial_merge_size ** 2
    out["vis.merger.ln_q"] = gain(Dv)
    out["vis.merger.fc1.weight"] = nrm(M, M, std=M ** -0.5); out["vis.merger.fc1.bias"] = nrm(M)
    out["vis.merger.fc2.weight"] = nrm(cfg.out_hidden, M, std=M ** -0.5); out["vis.merger.fc2.bias"] = nrm(cfg.out_hidden)
    D, kvd = cfg.hidden, cfg.kv_heads * cfg.head_dim
    out["llm.embed"] = nrm(cfg.vocab, D, std=1.0)
    out["llm.norm"] = gain(D)
    out["llm.lm_head"] = nrm(cfg.vocab, D, std=D ** -0.5)
    for l in range(cfg.layers):
        q = f"llm.{l}."
        out[q + "ln1"] = gain(D); out[q + "ln2"] = gain(D)
        out[q + "qkv.weight"] = nrm(D + 2 * kvd, D, std=D ** -0.5); out[q + "qkv.bias"] = nrm(D + 2 * kvd, std=0.1)
        out[q + "o.weight"] = nrm(D, D, std=D ** -0.5)
        out[q + "gate_up.weight"] = nrm(2 * cfg.mlp, D, std=D ** -0.5)
        out[q + "down.weight"] = nrm(D, cfg.mlp, std=cfg.mlp ** -0.5)
    return out


def synthetic_qwen_batch(cfg, batch: int, image_hw=(448, 448), text_len: int = 64, seed: int = 1, answer_id: int = 9454,
                         frames: int = 1):
    """BASELINE config 3: `batch` still images of image_hw (already multiples of 28) as processor-layout patches
    [batch * h/14 * w/14, 1176] fp32 (normalised noise), prompts of `text_len` text ids around one image-token run.
    frames > 1 (config 5): videos of `frames` temporal patches, grid (frames, h/14, w/14), video-token run."""
    g = torch.Generator().manual_seed(seed)
    gh, gw = image_hw[0] // cfg.patch_size, image_hw[1] // cfg.patch_size
    n_tok = frames * gh * gw // cfg.spatial_merge_size ** 2
    vis_id = cfg.image_token_id if frames == 1 else cfg.video_token_id
    patches = torch.randn(batch * frames * gh * gw, cfg.patch_dim, generator=g)
    prompts = []
    for b in range(batch):
        txt = torch.randint(0, min(cfg.image_token_id, cfg.video_token_id, cfg.vocab - 8), (text_len,), generator=g)
        pre = 14                                     # chat-template prefix length before <|vision_start|>
        prompts.append(torch.cat([txt[:pre], torch.full((n_tok,), vis_id), txt[pre:]]).tolist())
    return dict(pixel_patches=patches.pin_memory() if torch.cuda.is_available() else patches, grid_thw=[(frames, gh, gw)] * batch,
                prompts=prompts, answer_ids=[answer_id % cfg.vocab] * batch)
