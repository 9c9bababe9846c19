"""Generate tests/golden/*.pt from the REAL transformers modules (run in the build container; committed with its outputs).

  python tools/make_golden.py

For each case: seeded synthetic weights (oracle.make_synthetic_state_dict, bf16-representable), seeded inputs, and the
outputs of tests/hf_reference.hf_clipt5_forward (T5ForConditionalGeneration + CLIPVisionModel composed as the v3.0
wrapper did) in fp32 and under bf16 autocast. The micro case also stores the weights themselves so the fixture does not
depend on RNG stability; the larger cases store a checksum of the regenerated weights instead.
Also dumps golden vectors for the host-side pieces that survive in /root/reference (imported from there when present):
expand2square and t5_tokenizer_image_token.
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from oracle import clipt5_oracle as orc
import hf_reference as hf

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)


def sd_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().view(torch.uint8).numpy().tobytes())
    return h.hexdigest()


CASES = {
    # name: (config kwargs, batch, text_len, ragged, n_images, store_weights)
    "micro": (dict(image_size=28, vit_hidden=256, vit_heads=4, vit_mlp=256, vit_layers=2, d_model=128, n_heads=2, d_ff=256,
                   enc_layers=1, dec_layers=1, vocab=256), 3, 9, True, None, True),
    "tiny": (dict(), 4, 12, True, None, False),
    "tiny_shared_image": (dict(), 4, 10, False, 2, False),
    "mid": (dict(image_size=112, vit_hidden=1024, vit_heads=16, vit_mlp=1024, vit_layers=3, d_model=512, n_heads=8,
                 d_ff=1024, enc_layers=2, dec_layers=2, vocab=2048), 3, 24, True, None, False),
}


def main():
    for name, (kw, batch, L, ragged, n_images, store_w) in CASES.items():
        cfg = orc.ClipT5Config.tiny(**kw)
        label_ids = (37 % cfg.vocab, 1)
        sd = orc.make_synthetic_state_dict(cfg, seed=0, label_ids=label_ids)            # bf16
        inp = orc.make_synthetic_inputs(cfg, batch, L, seed=1, ragged=ragged, n_images=n_images, label_ids=label_ids)
        label_rows = orc.calibrate_lm_head(sd, cfg, inp)                                  # scores spread over (0,1)
        sd32 = {k: v.float() for k, v in sd.items()}
        mods = hf.build_hf_modules(cfg, sd32, dtype=torch.float32)
        r32 = hf.hf_clipt5_forward(cfg, mods, inp["pixels"], inp["input_ids"], inp["text_lens"], inp["labels"],
                                   inp["image_index"], return_all=True)
        mods16 = hf.build_hf_modules(cfg, sd32, dtype=torch.bfloat16)
        r16 = hf.hf_clipt5_forward(cfg, mods16, inp["pixels"], inp["input_ids"], inp["text_lens"], inp["labels"],
                                   inp["image_index"], autocast_bf16=True, return_all=True)
        lab = inp["labels"]
        logp32 = torch.log_softmax(r32["logits"], -1).gather(-1, lab[..., None])[..., 0]
        logp16 = torch.log_softmax(r16["logits"], -1).gather(-1, lab[..., None])[..., 0]
        blob = dict(config=kw, batch=batch, text_len=L, ragged=ragged, n_images=n_images, label_ids=label_ids,
                    seed_weights=0, seed_inputs=1, label_rows=label_rows, weights_sha256=sd_checksum(sd), inputs=inp,
                    hf_fp32=dict(scores=r32["scores"], logprobs=logp32, enc_absmean=float(r32["enc"].abs().mean()),
                                 enc_sample=r32["enc"][:, ::37, ::11].clone(), feats_sample=r32["feats"][:, ::5, ::13].clone()),
                    hf_bf16=dict(scores=r16["scores"], logprobs=logp16),
                    transformers_version=__import__("transformers").__version__, torch_version=torch.__version__)
        if store_w:
            blob["state_dict"] = sd
        torch.save(blob, os.path.join(GOLD, f"clipt5_{name}.pt"))
        print(name, "scores fp32", r32["scores"].tolist(), "bf16", r16["scores"].tolist())

    # ---- host-side vestiges of the reference (mm_utils.py:128-139,164-179), executed from /root/reference itself
    ref_mm = "/root/reference/t2v_metrics/models/vqascore_models/mm_utils.py"
    if os.path.exists(ref_mm):
        import importlib.util, types
        from PIL import Image
        import numpy as np
        # mm_utils imports `...constants` relatively and cv2; load it with a stub package context
        src = open(ref_mm).read()
        src = src.replace("from ...constants import", "from _ref_constants import")
        consts = types.ModuleType("_ref_constants")
        exec(open("/root/reference/t2v_metrics/constants.py").read(), consts.__dict__)
        sys.modules["_ref_constants"] = consts
        mod = types.ModuleType("_ref_mm_utils")
        try:
            exec(compile(src, ref_mm, "exec"), mod.__dict__)
        except ImportError as e:  # cv2 etc. missing: keep only the two pure functions
            import re
            keep = re.findall(r"(def expand2square.*?)(?=\ndef )", src, re.S) + re.findall(
                r"(def t5_tokenizer_image_token.*?)(?=\n\ndef )", src, re.S)
            mod.__dict__.update(dict(Image=Image, torch=torch, IMAGE_TOKEN_INDEX=consts.IMAGE_TOKEN_INDEX))
            exec("\n\n".join(keep), mod.__dict__)
        rng = np.random.RandomState(0)
        cases = []
        for (w, h) in [(40, 20), (20, 40), (33, 33), (1, 7)]:
            arr = rng.randint(0, 256, (h, w, 3), dtype=np.uint8)
            out = mod.expand2square(Image.fromarray(arr), (122, 116, 104))
            cases.append(dict(inp=torch.from_numpy(arr), out=torch.from_numpy(np.asarray(out).copy())))

        class FakeTok:  # deterministic stand-in with the T5 tokenizer's contract: ids + trailing </s>=1
            def __call__(self, chunk):
                ids = [3 + (ord(c) % 200) for c in chunk.split()] if False else [3 + (sum(map(ord, w)) % 997) for w in chunk.split()]
                return types.SimpleNamespace(input_ids=ids + [1])
        tok_cases = []
        for prompt in ["a b <image>\nc d e", "<image>", "no image here", "x <image> y <image> z"]:
            tok_cases.append(dict(prompt=prompt, ids=mod.t5_tokenizer_image_token(prompt, FakeTok())))
        torch.save(dict(expand2square=cases, t5_tokenizer_image_token=tok_cases), os.path.join(GOLD, "host_vestiges.pt"))
        print("host vestiges:", [c["ids"] for c in tok_cases])


if __name__ == "__main__":
    main()
