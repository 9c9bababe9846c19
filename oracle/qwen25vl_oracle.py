"""CPU oracle for the Qwen2.5-VL VQAScore hot path -- TEST INFRASTRUCTURE, NOT THE PRODUCT.

Plain-PyTorch restatement (fp32 or bf16-rounding-emulating) of what the reference runs for `qwen2.5-vl-*`:
t2v_metrics/models/vqascore_models/qwen2vl_model.py:169-301 (per-sample `generate(max_new_tokens=1, output_scores=True)`
then `softmax(scores[-1][0] / T)[answer_id]`, :160-167) on top of transformers 5.5.0
models/qwen2_5_vl/modeling_qwen2_5_vl.py (vision tower :345-518, mRoPE :545-669, GQA attention :672-759, decoder
:762-942, rope index :1024-1133, feature splice :1298-1307, lm_head :1515-1519). With `max_new_tokens=1` the generate
loop is a single prefill, so the oracle runs one forward and reads the last position (generation/utils.py:2487-2491).

Only tests/, __graft_entry__.smoke() and bench.py's cpu baseline may import this file. Pinned against the real
`Qwen2_5_VLForConditionalGeneration` on seeded tiny configs in tests/test_oracle_vs_hf.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from .clipt5_oracle import _Num


@dataclass
class Qwen25VLConfig:
    # vision tower (configuration_qwen2_5_vl.py:51-64; 7B checkpoint values)
    vit_depth: int = 32
    vit_hidden: int = 1280
    vit_heads: int = 16
    vit_mlp: int = 3420
    patch_size: int = 14
    temporal_patch_size: int = 2
    spatial_merge_size: int = 2
    window_size: int = 112
    fullatt_block_indexes: Tuple[int, ...] = (7, 15, 23, 31)
    out_hidden: int = 3584
    tokens_per_second: int = 2
    # language model
    hidden: int = 3584
    layers: int = 28
    heads: int = 28
    kv_heads: int = 4
    mlp: int = 18944
    vocab: int = 152064
    rms_eps: float = 1e-6
    rope_theta: float = 1e6
    mrope_section: Tuple[int, int, int] = (16, 24, 24)
    image_token_id: int = 151655
    video_token_id: int = 151656

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def patch_dim(self) -> int:
        return 3 * self.temporal_patch_size * self.patch_size ** 2

    @staticmethod
    def qwen25_vl_7b() -> "Qwen25VLConfig":
        return Qwen25VLConfig()

    @staticmethod
    def tiny(**kw) -> "Qwen25VLConfig":
        base = dict(vit_depth=4, vit_hidden=128, vit_heads=2, vit_mlp=172, fullatt_block_indexes=(1, 3), out_hidden=256,
                    hidden=256, layers=2, heads=4, kv_heads=2, mlp=512, vocab=640, mrope_section=(8, 12, 12),
                    image_token_id=600, video_token_id=601, window_size=56)
        base.update(kw)
        return Qwen25VLConfig(**base)


# ----------------------------------------------------------------------------------------------- host index logic
def vision_rot_pos_ids(grid_thw: Sequence[Sequence[int]], merge: int) -> torch.Tensor:
    """Qwen2_5_VisionTransformerPretrainedModel.rot_pos_emb (modeling_qwen2_5_vl.py:382-409): (h, w) index of every patch in
    merge-block order. Returns [sum(t*h*w), 2] int64."""
    out = []
    for t, h, w in grid_thw:
        hp = torch.arange(h).unsqueeze(1).expand(-1, w).reshape(h // merge, merge, w // merge, merge).permute(0, 2, 1, 3).flatten()
        wp = torch.arange(w).unsqueeze(0).expand(h, -1).reshape(h // merge, merge, w // merge, merge).permute(0, 2, 1, 3).flatten()
        out.append(torch.stack([hp, wp], dim=-1).repeat(t, 1))
    return torch.cat(out, dim=0)


def vision_window_index(grid_thw: Sequence[Sequence[int]], cfg: Qwen25VLConfig):
    """get_window_index (modeling_qwen2_5_vl.py:411-451). Returns (window_index over merged tokens, cu_window_seqlens over
    patches with consecutive duplicates removed, cu_seqlens of whole frames)."""
    merge, unit = cfg.spatial_merge_size, cfg.spatial_merge_size ** 2
    vw = cfg.window_size // merge // cfg.patch_size
    window_index, cu_win, base = [], [0], 0
    for t, h, w in grid_thw:
        gh, gw = h // merge, w // merge
        index = torch.arange(t * gh * gw).reshape(t, gh, gw)
        pad_h, pad_w = vw - gh % vw, vw - gw % vw
        nh, nw = (gh + pad_h) // vw, (gw + pad_w) // vw
        padded = F.pad(index, (0, pad_w, 0, pad_h), "constant", -100)
        padded = padded.reshape(t, nh, vw, nw, vw).permute(0, 1, 3, 2, 4).reshape(t, nh * nw, vw, vw)
        seqlens = (padded != -100).sum([2, 3]).reshape(-1)
        flat = padded.reshape(-1)
        window_index.append(flat[flat != -100] + base)
        cu_win.extend((seqlens.cumsum(0) * unit + cu_win[-1]).tolist())
        base += t * gh * gw
    cu_win_t = torch.unique_consecutive(torch.tensor(cu_win, dtype=torch.int32))
    frames = torch.tensor([h * w for t, h, w in grid_thw for _ in range(t)], dtype=torch.int32)
    cu_full = F.pad(frames.cumsum(0, dtype=torch.int32), (1, 0), value=0)
    return torch.cat(window_index), cu_win_t, cu_full


def mrope_position_ids(input_ids: torch.Tensor, grid_thw: Sequence[Sequence[int]], cfg: Qwen25VLConfig,
                       second_per_grid_ts: Optional[Sequence[float]] = None) -> torch.Tensor:
    """get_rope_index for ONE unpadded sequence with image / video tokens (modeling_qwen2_5_vl.py:1024-1133): text runs advance
    all three axes together, a vision run (image_token_id or video_token_id; `grid_thw` lists the grids in order of appearance)
    takes (t, h, w) grid indices offset by the running position and advances it by max(h, w) / merge. `second_per_grid_ts`: one
    entry per vision run in order (the reference's iterator is advanced by image runs too, :1113), default 1. Returns [3, L]."""
    ids = input_ids.tolist()
    merge = cfg.spatial_merge_size
    spg = iter(second_per_grid_ts) if second_per_grid_ts is not None else None
    kind = lambda tok: 1 if tok == cfg.image_token_id else (2 if tok == cfg.video_token_id else 0)   # mm_token_type_ids
    pos: List[torch.Tensor] = []
    cur, i, g = 0, 0, 0
    while i < len(ids):
        is_img = kind(ids[i])
        j = i
        while j < len(ids) and kind(ids[j]) == is_img:
            j += 1
        if not is_img:
            n = j - i
            pos.append(torch.arange(n).view(1, -1).expand(3, -1) + cur)
            cur += n
        else:
            t, h, w = grid_thw[g]
            g += 1
            gt, gh, gw = t, h // merge, w // merge
            assert j - i == gt * gh * gw, "image token run must match the grid"
            pw = torch.arange(cur, cur + gw).repeat(gh * gt)
            ph = torch.arange(cur, cur + gh).repeat_interleave(gw * gt)
            # transformers 5.5.0 get_vision_position_ids (:1018-1019): the temporal index is start_position * time_interval
            # with time_interval = tokens_per_second * second_per_grid (1 for images) -- also for still images (SURVEY 8c iii)
            interval = cfg.tokens_per_second * (int(next(spg)) if spg is not None else 1)
            pt = torch.full((gt * gh * gw,), cur * interval, dtype=torch.long)
            pos.append(torch.stack([pt, ph, pw], dim=0))
            cur += max(h, w) // merge
        i = j
    return torch.cat(pos, dim=1)


# ----------------------------------------------------------------------------------------------- numerics
def rms_norm(x, w, eps, num: _Num):
    """Qwen2_5_VLRMSNorm (modeling_qwen2_5_vl.py:66-71): fp32 normalise, cast to the input dtype, then scale."""
    xf = x.float()
    xn = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return num.r(num.r(w.float()) * num.r(xn))


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def vision_tower(sd, pixel_patches: torch.Tensor, grid_thw, cfg: Qwen25VLConfig, mode="fp32", prefix="model.visual."):
    """Qwen2_5_VisionTransformerPretrainedModel.forward (:455-518). pixel_patches [sum(t*h*w), 1176] -> merged features
    [sum(t*h*w)/4, out_hidden] in ORIGINAL (un-windowed) order."""
    num = _Num(mode)
    g = lambda k: sd[prefix + k].float()
    D, H = cfg.vit_hidden, cfg.vit_heads
    hd = D // H
    unit = cfg.spatial_merge_size ** 2
    x = num.linear(pixel_patches.float(), g("patch_embed.proj.weight").reshape(D, -1))          # Conv3d == GEMM, no bias
    L = x.shape[0]
    pos_ids = vision_rot_pos_ids(grid_thw, cfg.spatial_merge_size)
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, hd // 2, 2, dtype=torch.float) / (hd // 2)))
    max_grid = max(max(h, w) for _, h, w in grid_thw)
    freqs_full = torch.outer(torch.arange(max_grid, dtype=torch.float), inv_freq)                # [grid, hd/4]
    rot = freqs_full[pos_ids].flatten(1)                                                         # [L, hd/2]
    widx, cu_win, cu_full = vision_window_index(grid_thw, cfg)
    x = x.reshape(L // unit, unit, -1)[widx].reshape(L, -1)
    rot = rot.reshape(L // unit, unit, -1)[widx].reshape(L, -1)
    emb = torch.cat((rot, rot), dim=-1)
    cos, sin = emb.cos()[:, None, :], emb.sin()[:, None, :]                                       # [L, 1, hd]
    for l in range(cfg.vit_depth):
        p = f"blocks.{l}."
        cu = cu_full if l in cfg.fullatt_block_indexes else cu_win
        y = rms_norm(x, g(p + "norm1.weight"), 1e-6, num)
        qkv = num.linear(y, g(p + "attn.qkv.weight"), g(p + "attn.qkv.bias")).reshape(L, 3, H, hd)
        q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
        q = num.r(q.float() * cos + rotate_half(q.float()) * sin)                                # apply_rotary_pos_emb_vision
        k = num.r(k.float() * cos + rotate_half(k.float()) * sin)
        outs = []
        for s, e in zip(cu[:-1].tolist(), cu[1:].tolist()):
            qq, kk, vv = (t[s:e].transpose(0, 1) for t in (q, k, v))                            # [H, n, hd]
            att = num.r(torch.matmul(qq, kk.transpose(1, 2)) * hd ** -0.5)
            att = num.r(torch.softmax(att.float(), dim=-1))
            outs.append(num.r(torch.matmul(att, vv)).transpose(0, 1).reshape(e - s, D))
        o = torch.cat(outs, dim=0)
        x = num.r(x + num.linear(o, g(p + "attn.proj.weight"), g(p + "attn.proj.bias")))
        y = rms_norm(x, g(p + "norm2.weight"), 1e-6, num)
        gate = num.linear(y, g(p + "mlp.gate_proj.weight"), g(p + "mlp.gate_proj.bias"))
        up = num.linear(y, g(p + "mlp.up_proj.weight"), g(p + "mlp.up_proj.bias"))
        x = num.r(x + num.linear(num.r(num.r(F.silu(gate)) * up), g(p + "mlp.down_proj.weight"), g(p + "mlp.down_proj.bias")))
    y = rms_norm(x, g("merger.ln_q.weight"), 1e-6, num).reshape(L // unit, unit * D)
    y = num.linear(y, g("merger.mlp.0.weight"), g("merger.mlp.0.bias"))
    y = num.linear(num.r(F.gelu(y)), g("merger.mlp.2.weight"), g("merger.mlp.2.bias"))
    return y[torch.argsort(widx)]


def text_last_logits(sd, embeds: torch.Tensor, position_ids: torch.Tensor, cfg: Qwen25VLConfig, mode="fp32",
                     prefix="model.language_model.") -> torch.Tensor:
    """Decoder prefill for ONE sequence (the reference scores one sample at a time, qwen2vl_model.py:190) and lm_head on the
    last position only. embeds [L, hidden], position_ids [3, L] -> fp32 logits [vocab]."""
    num = _Num(mode)
    g = lambda k: sd[prefix + k].float()
    L = embeds.shape[0]
    Hq, Hkv, hd = cfg.heads, cfg.kv_heads, cfg.head_dim
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.float) / hd))
    freqs = position_ids[:, :, None].float() * inv_freq[None, None, :]                           # [3, L, hd/2]
    emb = torch.cat((freqs, freqs), dim=-1)
    cos3, sin3 = num.r(emb.cos()), num.r(emb.sin())                                               # cast to the activation dtype
    sec = list(cfg.mrope_section) * 2
    cos = torch.cat([m[i % 3] for i, m in enumerate(cos3.split(sec, dim=-1))], dim=-1)[:, None, :]   # [L, 1, hd]
    sin = torch.cat([m[i % 3] for i, m in enumerate(sin3.split(sec, dim=-1))], dim=-1)[:, None, :]
    causal = torch.tril(torch.ones(L, L, dtype=torch.bool))
    neg = torch.finfo(torch.bfloat16 if num.bf16 else torch.float32).min
    x = embeds
    for l in range(cfg.layers):
        p = f"layers.{l}."
        y = rms_norm(x, g(p + "input_layernorm.weight"), cfg.rms_eps, num)
        q = num.linear(y, g(p + "self_attn.q_proj.weight"), g(p + "self_attn.q_proj.bias")).view(L, Hq, hd)
        k = num.linear(y, g(p + "self_attn.k_proj.weight"), g(p + "self_attn.k_proj.bias")).view(L, Hkv, hd)
        v = num.linear(y, g(p + "self_attn.v_proj.weight"), g(p + "self_attn.v_proj.bias")).view(L, Hkv, hd)
        q = num.r(q * cos + rotate_half(q) * sin)
        k = num.r(k * cos + rotate_half(k) * sin)
        rep = Hq // Hkv
        kk = k.repeat_interleave(rep, dim=1).transpose(0, 1)                                     # [Hq, L, hd]
        vv = v.repeat_interleave(rep, dim=1).transpose(0, 1)
        att = num.r(torch.matmul(q.transpose(0, 1), kk.transpose(1, 2)) * hd ** -0.5)
        att = num.r(att + torch.where(causal, 0.0, neg)[None])
        att = num.r(torch.softmax(att.float(), dim=-1))
        o = num.r(torch.matmul(att, vv)).transpose(0, 1).reshape(L, Hq * hd)
        x = num.r(x + num.linear(o, g(p + "self_attn.o_proj.weight")))
        y = rms_norm(x, g(p + "post_attention_layernorm.weight"), cfg.rms_eps, num)
        gate = num.linear(y, g(p + "mlp.gate_proj.weight"))
        up = num.linear(y, g(p + "mlp.up_proj.weight"))
        x = num.r(x + num.linear(num.r(num.r(F.silu(gate)) * up), g(p + "mlp.down_proj.weight")))
    h = rms_norm(x[-1:], g("norm.weight"), cfg.rms_eps, num)
    return num.linear(h, sd["lm_head.weight"])[0].float()


def answer_probability(logits: torch.Tensor, answer_id: int, temperature: float = 1.0, prompt_ids: Optional[torch.Tensor] = None,
                       repetition_penalty: float = 1.0) -> torch.Tensor:
    """qwen2vl_model.py:160-167 on `outputs.scores[-1][0]`: fp32 scores AFTER the logits processors
    (generation/utils.py:2762-2770). The only processor that can be active for this call is the repetition penalty of the
    checkpoint's generation_config: logits of ids present in the prompt are divided (if > 0) or multiplied
    (if < 0) by the penalty."""
    s = logits.float().clone()
    if repetition_penalty != 1.0 and prompt_ids is not None:
        idx = torch.unique(prompt_ids)
        v = s[idx]
        s[idx] = torch.where(v < 0, v * repetition_penalty, v / repetition_penalty)
    return torch.softmax(s / temperature, dim=-1)[answer_id]


@torch.no_grad()
def qwen25vl_score(sd: Dict[str, torch.Tensor], cfg: Qwen25VLConfig, pixel_patches: torch.Tensor, grid_thw: Sequence[Sequence[int]],
                   input_ids: List[torch.Tensor], answer_ids: Sequence[int], image_of_sample: Optional[Sequence[int]] = None,
                   mode: str = "fp32", temperature: float = 1.0, repetition_penalty: float = 1.0, return_all: bool = False,
                   second_per_grid_ts: Optional[Sequence[float]] = None):
    """pixel_patches: all images' (or videos') patches concatenated [sum P, patch_dim]; grid_thw one (t,h,w) per image / video;
    input_ids: one 1-D id tensor per sample containing a run of image_token_id (or video_token_id) for its visual input -- the
    model treats both the same way: same tower, masked_scatter at the run (:1298-1322); answer_ids[b]: the single answer token
    id; second_per_grid_ts: one per image / video (temporal patch duration in seconds, videos only; default 1).
    Returns probabilities [B] (fp32)."""
    num = _Num(mode)
    feats = vision_tower(sd, pixel_patches, grid_thw, cfg, mode)
    unit = cfg.spatial_merge_size ** 2
    offs = [0]
    for t, h, w in grid_thw:
        offs.append(offs[-1] + t * h * w // unit)
    embed_w = num.r(sd["model.language_model.embed_tokens.weight"].float())
    probs, all_logits = [], []
    for b, ids in enumerate(input_ids):
        img = image_of_sample[b] if image_of_sample is not None else b
        e = embed_w[ids].clone()
        mask = (ids == cfg.image_token_id) | (ids == cfg.video_token_id)
        e[mask] = num.r(feats[offs[img]:offs[img + 1]])                                           # masked_scatter (:1301-1307)
        pos = mrope_position_ids(ids, [grid_thw[img]], cfg, None if second_per_grid_ts is None else [second_per_grid_ts[img]])
        logits = text_last_logits(sd, e, pos, cfg, mode)
        all_logits.append(logits)
        probs.append(answer_probability(logits, int(answer_ids[b]), temperature, ids, repetition_penalty))
    out = torch.stack(probs)
    if return_all:
        return dict(scores=out, logits=torch.stack(all_logits), feats=feats)
    return out


# ----------------------------------------------------------------------------------------------- synthetic weights / inputs
def make_synthetic_state_dict(cfg: Qwen25VLConfig, seed: int = 0, dtype=torch.bfloat16, gen_device="cpu",
                              pool: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    # gen_device="cpu" is the seeded sequence behind the committed goldens; the full-width GPU tests draw 8 B parameters on the device
    gd = torch.device(gen_device)
    g = torch.Generator(device=gd).manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    cursor = [0]

    def randn(*shape):
        if pool is None:
            return torch.randn(*shape, generator=g, device=gd)
        # timing-only weights (bench.py's CPU baseline): values cycled out of a small pre-drawn pool instead of 8 B fresh draws
        n = 1
        for d in shape:
            n *= int(d)
        P = pool.numel()
        off = cursor[0] % P
        cursor[0] += n + 7919
        out = torch.empty(n, dtype=pool.dtype)
        pos = min(P - off, n)
        out[:pos] = pool[off:off + pos]
        while pos < n:
            m = min(P, n - pos)
            out[pos:pos + m] = pool[:m]
            pos += m
        return out.view(*shape).to(gd)

    def nrm(name, *shape, std=0.02):
        sd[name] = (randn(*shape) * std).to(dtype)

    def gain(name, n):
        sd[name] = (1.0 + 0.1 * randn(n)).to(dtype)

    D, v = cfg.vit_hidden, "model.visual."
    nrm(v + "patch_embed.proj.weight", D, 3, cfg.temporal_patch_size, cfg.patch_size, cfg.patch_size, std=cfg.patch_dim ** -0.5)
    for l in range(cfg.vit_depth):
        p = v + f"blocks.{l}."
        gain(p + "norm1.weight", D); gain(p + "norm2.weight", D)
        nrm(p + "attn.qkv.weight", 3 * D, D, std=D ** -0.5); nrm(p + "attn.qkv.bias", 3 * D)
        nrm(p + "attn.proj.weight", D, D, std=D ** -0.5); nrm(p + "attn.proj.bias", D)
        for n_ in ("gate_proj", "up_proj"):
            nrm(p + f"mlp.{n_}.weight", cfg.vit_mlp, D, std=D ** -0.5); nrm(p + f"mlp.{n_}.bias", cfg.vit_mlp)
        nrm(p + "mlp.down_proj.weight", D, cfg.vit_mlp, std=cfg.vit_mlp ** -0.5); nrm(p + "mlp.down_proj.bias", D)
    M = D * cfg.spatial_merge_size ** 2
    gain(v + "merger.ln_q.weight", D)
    nrm(v + "merger.mlp.0.weight", M, M, std=M ** -0.5); nrm(v + "merger.mlp.0.bias", M)
    nrm(v + "merger.mlp.2.weight", cfg.out_hidden, M, std=M ** -0.5); nrm(v + "merger.mlp.2.bias", cfg.out_hidden)
    Hd, t = cfg.hidden, "model.language_model."
    nrm(t + "embed_tokens.weight", cfg.vocab, Hd, std=1.0)
    kvd = cfg.kv_heads * cfg.head_dim
    for l in range(cfg.layers):
        p = t + f"layers.{l}."
        gain(p + "input_layernorm.weight", Hd); gain(p + "post_attention_layernorm.weight", Hd)
        nrm(p + "self_attn.q_proj.weight", Hd, Hd, std=Hd ** -0.5); nrm(p + "self_attn.q_proj.bias", Hd, std=0.1)
        nrm(p + "self_attn.k_proj.weight", kvd, Hd, std=Hd ** -0.5); nrm(p + "self_attn.k_proj.bias", kvd, std=0.1)
        nrm(p + "self_attn.v_proj.weight", kvd, Hd, std=Hd ** -0.5); nrm(p + "self_attn.v_proj.bias", kvd, std=0.1)
        nrm(p + "self_attn.o_proj.weight", Hd, Hd, std=Hd ** -0.5)
        nrm(p + "mlp.gate_proj.weight", cfg.mlp, Hd, std=Hd ** -0.5)
        nrm(p + "mlp.up_proj.weight", cfg.mlp, Hd, std=Hd ** -0.5)
        nrm(p + "mlp.down_proj.weight", Hd, cfg.mlp, std=cfg.mlp ** -0.5)
    gain(t + "norm.weight", Hd)
    nrm("lm_head.weight", cfg.vocab, Hd, std=Hd ** -0.5)
    return sd


def make_synthetic_inputs(cfg: Qwen25VLConfig, batch: int, image_hw: Tuple[int, int] = (56, 56), text_len: int = 12, seed: int = 1,
                          ragged: bool = False, n_images: Optional[int] = None, answer_id: int = 9, frames: int = 1):
    """Patches as the Qwen2-VL image processor lays them out (image_processing_qwen2_vl.py:191-220): one still image ->
    grid (1, H/14, W/14), each row = one 2x14x14x3 patch (the frame duplicated along the temporal axis), in merge-block
    order. Ids: `text_len` text ids with one run of image_token_id of length h*w/4 somewhere inside."""
    g = torch.Generator().manual_seed(seed)
    ni = n_images or batch
    gh, gw = image_hw[0] // cfg.patch_size, image_hw[1] // cfg.patch_size
    # frames > 1: a video of `frames` temporal patches (grid (frames, gh, gw), video_token_id run; image_processing / video
    # processing lay patches out frame-major in the same merge-block order)
    grid = [(frames, gh, gw)] * ni
    n_tok = frames * gh * gw // cfg.spatial_merge_size ** 2
    vis_id = cfg.image_token_id if frames == 1 else cfg.video_token_id
    patches = torch.randn(ni * frames * gh * gw, cfg.patch_dim, generator=g)
    ids, img_of = [], []
    for b in range(batch):
        n = text_len if not ragged else int(torch.randint(max(4, text_len // 2), text_len + 1, (1,), generator=g))
        pre = int(torch.randint(1, n - 1, (1,), generator=g))
        txt = torch.randint(0, min(cfg.image_token_id, cfg.video_token_id, cfg.vocab - 8), (n,), generator=g)
        ids.append(torch.cat([txt[:pre], torch.full((n_tok,), vis_id), txt[pre:]]))
        img_of.append(b % ni)
    return dict(pixel_patches=patches, grid_thw=grid, input_ids=ids, answer_ids=[answer_id] * batch,
                image_of_sample=None if ni == batch else img_of)
