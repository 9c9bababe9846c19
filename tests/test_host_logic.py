"""Host-side logic of the drop-in boundary: tokenizer splice, expand2square, preprocessing, registry/API surface, sharding.
CPU only."""
import os
import types

import numpy as np
import pytest
import torch
from PIL import Image

import t2v_metrics_b200 as t2v
from t2v_metrics_b200.models.vqascore_models import mm_utils
from t2v_metrics_b200.models.vqascore_models.clip_t5_model import format_question, CLIP_T5_MODELS, CLIPT5Model
from t2v_metrics_b200.parallel import shard_bounds
from oracle import clipt5_oracle as orc


class FakeTok:
    """Same contract as the generator in tools/make_golden.py: per-word ids + trailing </s> = 1."""
    pad_token_id = 0

    def __call__(self, chunk):
        return types.SimpleNamespace(input_ids=[3 + (sum(map(ord, w)) % 997) for w in chunk.split()] + [1])


def test_golden_vestiges_from_reference(golden_dir):
    """expand2square / t5_tokenizer_image_token outputs recorded by running /root/reference's own mm_utils.py."""
    blob = torch.load(os.path.join(golden_dir, "host_vestiges.pt"), weights_only=False)
    for case in blob["expand2square"]:
        img = Image.fromarray(case["inp"].numpy())
        for fn in (mm_utils.expand2square, orc.expand2square):
            out = np.asarray(fn(img, (122, 116, 104)))
            assert np.array_equal(out, case["out"].numpy())
    for case in blob["t5_tokenizer_image_token"]:
        assert mm_utils.t5_tokenizer_image_token(case["prompt"], FakeTok()) == case["ids"]
        assert orc.t5_tokenizer_image_token(case["prompt"], lambda c: FakeTok()(c).input_ids) == case["ids"]


def test_tokenizer_chunk_cache_is_exact_and_saves_calls():
    """the memo returns exactly what the uncached call returns and tokenises each distinct chunk once."""
    class Counting(FakeTok):
        calls = 0

        def __call__(self, chunk):
            Counting.calls += 1
            return super().__call__(chunk)

    tok, cache = Counting(), {}
    prompts = [f"sys USER: <image>\nDoes this figure show \"{c}\"? ASSISTANT: " for c in ("a dog", "two cats", "a dog", "a dog")]
    plain = [mm_utils.t5_tokenizer_image_token(p, FakeTok()) for p in prompts]
    cached = [mm_utils.t5_tokenizer_image_token(p, tok, chunk_cache=cache) for p in prompts]
    assert cached == plain
    assert Counting.calls == 3            # the shared system chunk + two distinct caption chunks
    from t2v_metrics_b200.models.vqascore_models.qwen_utils import build_prompt_ids

    class QTok:
        calls = 0

        def encode(self, s, add_special_tokens=False):
            QTok.calls += 1
            return [5 + (ord(c) % 89) for c in s]

    qc = {}
    a = [build_prompt_ids(QTok(), q, 4, 600, qc) for q in ("x?", "y?", "x?")]
    b = [build_prompt_ids(QTok(), q, 4, 600) for q in ("x?", "y?", "x?")]
    assert a == b and a[0].count(600) == 4


def test_tokenizer_each_chunk_keeps_its_eos():
    ids = mm_utils.t5_tokenizer_image_token("a <image>\nb", FakeTok())
    assert ids.count(-200) == 1 and ids[ids.index(-200) - 1] == 1 and ids[-1] == 1


def test_prompt_format_matches_reference_constants():
    q = format_question('Does this figure show "a dog"? Please answer yes or no.')
    assert q.startswith("A chat between a curious user") and " USER: <image>\n" in q and q.endswith(" ASSISTANT: ")
    assert q == orc.format_question('Does this figure show "a dog"? Please answer yes or no.')


def test_clip_preprocess_matches_hf_processor():
    from transformers import CLIPImageProcessor
    rng = np.random.RandomState(0)
    proc = CLIPImageProcessor(size={"shortest_edge": 336}, crop_size={"height": 336, "width": 336}, do_resize=True,
                              do_center_crop=True, do_normalize=True, image_mean=list(mm_utils.CLIP_IMAGE_MEAN),
                              image_std=list(mm_utils.CLIP_IMAGE_STD), resample=3, do_convert_rgb=True)
    for (w, h) in [(512, 512), (640, 400), (300, 500)]:
        img = Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8))
        sq = mm_utils.expand2square(img, tuple(int(x * 255) for x in mm_utils.CLIP_IMAGE_MEAN))
        ref = torch.from_numpy(np.asarray(proc.preprocess(sq, return_tensors="np")["pixel_values"][0]))
        got = mm_utils.clip_preprocess(img, 336, pad=True)
        assert got.shape == (3, 336, 336)
        assert float((got - ref).abs().max()) < 2e-2          # PIL vs processor resampling paths, <= 1 grey level
        assert float((got - orc.clip_preprocess(img, 336)).abs().max()) == 0.0


def test_registry_surface_matches_reference():
    assert t2v.list_all_models() == ["clip-flant5-xxl", "clip-flant5-xl", "qwen2.5-vl-7b"]
    assert CLIPT5Model.video_mode == "concat" and CLIPT5Model.allows_image
    assert set(CLIP_T5_MODELS["clip-flant5-xxl"]) >= {"tokenizer", "model"}
    with pytest.raises(NotImplementedError):
        t2v.get_score_model("not-a-model")
    cfg = CLIP_T5_MODELS["clip-flant5-xxl"]["config"]()
    assert (cfg.d_model, cfg.n_heads, cfg.d_ff, cfg.vocab, cfg.num_patches) == (4096, 64, 10240, 32128, 576)


def test_engine_refuses_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from t2v_metrics_b200.engine import ClipT5Engine
    from t2v_metrics_b200.config import ClipT5Config
    with pytest.raises(RuntimeError):
        ClipT5Engine(ClipT5Config())


@pytest.mark.parametrize("n,world", [(10000, 8), (7, 8), (64, 1), (65, 2), (0, 4)])
def test_shard_bounds_cover_all_pairs_once(n, world):
    seen = []
    for r in range(world):
        s, e, per = shard_bounds(n, world, r)
        assert 0 <= s <= e <= n and e - s <= per
        seen.extend(range(s, e))
    assert seen == list(range(n))


def test_qwen_patch_layout_matches_hf_processor():
    """a19: smart_resize + patch flattening vs transformers' Qwen2VLImageProcessor (do_resize=False on a pre-sized image, as the
    reference calls it after qwen_vl_utils resized the image: qwen2vl_model.py:208-216)."""
    from transformers.models.qwen2_vl.image_processing_qwen2_vl import Qwen2VLImageProcessor, smart_resize as hf_smart_resize
    from t2v_metrics_b200.models.vqascore_models.qwen_utils import smart_resize, qwen_image_to_patches
    for hw in [(448, 448), (500, 333), (30, 4000), (2000, 3000), (57, 57)]:
        assert smart_resize(*hw) == hf_smart_resize(*hw)
    rng = np.random.RandomState(0)
    proc = Qwen2VLImageProcessor()
    for (h, w) in [(448, 448), (84, 140)]:
        img = Image.fromarray(rng.randint(0, 256, (h, w, 3), dtype=np.uint8))
        ref = proc(images=[img], do_resize=False, return_tensors="pt")
        got, grid = qwen_image_to_patches(img)
        assert list(grid) == ref["image_grid_thw"][0].tolist()
        assert float((got - ref["pixel_values"]).abs().max()) < 1e-5


def test_prompt_constants_equal_the_reference_module():
    """The prompt constants are part of the model's training recipe: compare with the reference's own module when it is mounted
    (build container); on the GPU box the reference is absent and the values are pinned literally."""
    import importlib.util
    from t2v_metrics_b200 import constants as c
    assert c.IMAGE_TOKEN_INDEX == -200 and c.IGNORE_INDEX == -100 and c.DEFAULT_IMAGE_TOKEN == "<image>" and c.CONTEXT_LEN == 2048
    assert c.SYSTEM_MSG.startswith("A chat between a curious user") and c.SYSTEM_MSG.endswith("to the user's questions.")
    ref = "/root/reference/t2v_metrics/constants.py"
    if not os.path.exists(ref):
        return
    spec = importlib.util.spec_from_file_location("ref_constants", ref)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    for k in ("HF_CACHE_DIR", "CONTEXT_LEN", "SYSTEM_MSG", "IGNORE_INDEX", "IMAGE_TOKEN_INDEX", "DEFAULT_IMAGE_TOKEN"):
        assert getattr(c, k) == getattr(m, k), k


def test_image_loader_cases(tmp_path):
    """.npy = BGR array reversed to RGB; everything else through PIL, always 3-channel RGB (reference model.py:10-14)."""
    from t2v_metrics_b200.models.model import image_loader
    a = (np.arange(4 * 5 * 3) % 255).astype(np.uint8).reshape(4, 5, 3)
    np.save(tmp_path / "x.npy", a)
    Image.fromarray(a).save(tmp_path / "x.png")
    Image.fromarray(a[:, :, 0]).save(tmp_path / "g.png")
    assert (np.asarray(image_loader(str(tmp_path / "x.npy"))) == a[:, :, ::-1]).all()
    assert (np.asarray(image_loader(str(tmp_path / "x.png"))) == a).all()
    assert np.asarray(image_loader(str(tmp_path / "g.png"))).shape == (4, 5, 3)
