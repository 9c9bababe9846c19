"""Generate tests/golden/qwen_tiny.pt from the REAL transformers Qwen2_5_VLForConditionalGeneration (run in the build container; the
fixture is committed with this script). Seeded synthetic weights (bf16-representable) of a tiny config with head_dim 128, two cases:
  * images: 4 prompts over 2 images of 84x56 pixels (grid 1x6x4), ragged text lengths
  * video:  3 prompts, each its own 2-temporal-patch clip of 56x84 (grid 2x4x6), second_per_grid 2.0
For each prompt: the fp32 last-position logits of the HF model, the answer probability (softmax(logits)[answer]), and the same after HF's
RepetitionPenaltyLogitsProcessor(1.3) + temperature 0.5 -- what `generate(max_new_tokens=1, output_scores=True)` returns in the reference
(t2v_metrics/models/vqascore_models/qwen2vl_model.py:160-167, 222-230).

  python tools/make_golden_qwen.py
"""
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from transformers.generation.logits_process import RepetitionPenaltyLogitsProcessor

from oracle import qwen25vl_oracle as qo
from test_qwen_host import hf_model, TINY

GOLD = os.path.join(ROOT, "tests", "golden")


def sd_checksum(sd):
    import hashlib
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().view(torch.uint8).numpy().tobytes())
    return h.hexdigest()


@torch.no_grad()
def hf_case(m, cfg, inp, video, spg):
    P = inp["grid_thw"][0][0] * inp["grid_thw"][0][1] * inp["grid_thw"][0][2]
    logits = []
    for b, ids in enumerate(inp["input_ids"]):
        img = inp["image_of_sample"][b] if inp["image_of_sample"] is not None else b
        px = inp["pixel_patches"][img * P:(img + 1) * P]
        grid = torch.tensor([list(inp["grid_thw"][img])])
        common = dict(input_ids=ids[None], attention_mask=torch.ones(1, len(ids), dtype=torch.long))
        if video:
            out = m(pixel_values_videos=px, video_grid_thw=grid, second_per_grid_ts=torch.tensor([spg[img]]),
                    mm_token_type_ids=torch.where(ids == cfg.video_token_id, 2, 0)[None], **common)
        else:
            out = m(pixel_values=px, image_grid_thw=grid, mm_token_type_ids=(ids == cfg.image_token_id).long()[None], **common)
        logits.append(out.logits[0, -1].float())
    logits = torch.stack(logits)
    ans = torch.tensor(inp["answer_ids"])
    probs = torch.softmax(logits, -1)[torch.arange(len(ans)), ans]
    proc = RepetitionPenaltyLogitsProcessor(1.3)
    pen = torch.stack([torch.softmax(proc(ids[None], logits[b][None].clone())[0] / 0.5, -1)[ans[b]] for b, ids in enumerate(inp["input_ids"])])
    return dict(logits=logits, probs=probs, probs_penalty_1p3_T_0p5=pen)


def main():
    cfg = qo.Qwen25VLConfig.tiny(**TINY)
    sd = qo.make_synthetic_state_dict(cfg, seed=0)          # bf16
    m = hf_model(cfg)
    m.load_state_dict({k: v.float() for k, v in sd.items()})
    # the weights are regenerated from the seed by the tests (same image, same torch); the checksum pins them
    blob = dict(config=dataclasses.asdict(cfg), weights_seed=0, weights_sha256=sd_checksum(sd), cases={})
    img = qo.make_synthetic_inputs(cfg, 4, (84, 56), 12, seed=1, ragged=True, n_images=2)
    img["answer_ids"][0] = int(img["input_ids"][0][-1])       # an answer id that occurs in its own prompt: the penalty moves the label logit too
    blob["cases"]["images"] = dict(inputs=img, second_per_grid_ts=None, hf=hf_case(m, cfg, img, False, None))
    vid = qo.make_synthetic_inputs(cfg, 3, (56, 84), 10, seed=2, ragged=True, frames=2)
    spg = [2.0, 2.0, 2.0]
    blob["cases"]["video"] = dict(inputs=vid, second_per_grid_ts=spg, hf=hf_case(m, cfg, vid, True, spg))
    path = os.path.join(GOLD, "qwen_tiny.pt")
    torch.save(blob, path)
    for name, c in blob["cases"].items():
        print(name, "probs", [round(float(x), 6) for x in c["hf"]["probs"]], "penalised", [round(float(x), 6) for x in c["hf"]["probs_penalty_1p3_T_0p5"]])
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
