"""Pin the oracle (oracle/clipt5_oracle.py) to the real transformers modules and to the committed golden vectors.
CPU only. The reference repository pins no numeric outputs itself."""
import dataclasses
import os

import pytest
import torch

from oracle import clipt5_oracle as orc
import hf_reference as hf


def _case(kw, batch=3, L=10, ragged=True, n_images=None, seed=0):
    cfg = orc.ClipT5Config.tiny(**kw)
    label_ids = (37 % cfg.vocab, 1)
    sd = orc.make_synthetic_state_dict(cfg, seed=seed, label_ids=label_ids)
    inp = orc.make_synthetic_inputs(cfg, batch, L, seed=seed + 1, ragged=ragged, n_images=n_images, label_ids=label_ids)
    return cfg, sd, inp


@pytest.mark.parametrize("kw", [dict(), dict(image_size=28, vit_layers=2, enc_layers=1, dec_layers=3, n_heads=2, d_model=128)])
def test_oracle_fp32_matches_hf_stage_by_stage(kw):
    cfg, sd, inp = _case(kw)
    sd32 = {k: v.float() for k, v in sd.items()}
    mods = hf.build_hf_modules(cfg, sd32)
    ref = hf.hf_clipt5_forward(cfg, mods, inp["pixels"], inp["input_ids"], inp["text_lens"], inp["labels"], return_all=True)
    got = orc.clipt5_score(sd, cfg, inp["pixels"], inp["input_ids"], inp["text_lens"], inp["labels"], mode="fp32", return_all=True)
    for k in ("feats", "proj", "embeds", "enc", "logits"):
        scale = float(ref[k].abs().max())
        assert float((ref[k] - got[k]).abs().max()) <= 2e-5 * max(scale, 1.0), k
    assert torch.equal(ref["mask"], got["mask"])
    assert float((ref["scores"] - got["scores"]).abs().max()) < 1e-5


def test_oracle_bf16_mode_tracks_hf_autocast():
    """bf16 emulation is not bit-identical to HF autocast (different fp32 accumulation orders), but both must sit within the
    same distance of the fp32 result."""
    cfg, sd, inp = _case(dict())
    orc.calibrate_lm_head(sd, cfg, inp)
    sd32 = {k: v.float() for k, v in sd.items()}
    ref32 = orc.clipt5_score(sd, cfg, inp["pixels"], inp["input_ids"], inp["text_lens"], inp["labels"], mode="fp32", return_all=True)
    mods16 = hf.build_hf_modules(cfg, sd32, dtype=torch.bfloat16)
    hf16 = hf.hf_clipt5_forward(cfg, mods16, inp["pixels"], inp["input_ids"], inp["text_lens"], inp["labels"], autocast_bf16=True,
                                return_all=True)
    or16 = orc.clipt5_score(sd, cfg, inp["pixels"], inp["input_ids"], inp["text_lens"], inp["labels"], mode="bf16", return_all=True)
    e_hf = float((hf16["logits"] - ref32["logits"]).abs().max())
    e_or = float((or16["logits"] - ref32["logits"]).abs().max())
    assert e_or < 3 * e_hf + 0.05, (e_or, e_hf)
    assert float((or16["scores"] - hf16["scores"]).abs().max()) < 2e-2


@pytest.mark.parametrize("name", ["micro", "tiny", "tiny_shared_image", "mid"])
def test_oracle_reproduces_golden_hf_vectors(name, golden_dir):
    blob = torch.load(os.path.join(golden_dir, f"clipt5_{name}.pt"), weights_only=False)
    cfg = orc.ClipT5Config.tiny(**blob["config"])
    if "state_dict" in blob:
        sd = blob["state_dict"]
    else:
        sd = orc.make_synthetic_state_dict(cfg, seed=blob["seed_weights"], label_ids=blob["label_ids"])
        lm = sd["lm_head.weight"].float()
        for t, row in zip(blob["label_ids"], blob["label_rows"]):
            lm[t] = row
        sd["lm_head.weight"] = lm.to(torch.bfloat16)
    inp = blob["inputs"]
    got = orc.clipt5_score(sd, cfg, inp["pixels"], inp["input_ids"], inp["text_lens"], inp["labels"], inp["image_index"],
                           mode="fp32", return_all=True)
    gold = blob["hf_fp32"]
    assert float((got["scores"] - gold["scores"]).abs().max()) < 2e-5
    assert float((got["logprobs"] - gold["logprobs"]).abs().max()) < 2e-4
    assert float((got["enc"][:, ::37, ::11] - gold["enc_sample"]).abs().max()) < 2e-4
    assert float((got["feats"][:, ::5, ::13] - gold["feats_sample"]).abs().max()) < 2e-4
    # the reference's own bf16-autocast run differs from its fp32 run by more than 1e-3 already on these shallow models
    blob["hf_bf16"]["scores"]


def test_relative_position_bucket_matches_hf():
    from transformers.models.t5.modeling_t5 import T5Attention
    rel = torch.arange(-700, 701)[None, :]
    for bidir in (True, False):
        a = T5Attention._relative_position_bucket(rel, bidirectional=bidir, num_buckets=32, max_distance=128)
        b = orc.relative_position_bucket(rel, bidir, 32, 128)
        assert torch.equal(a, b)
