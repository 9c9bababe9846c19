"""Qwen2.5-VL VQAScore plugin backed by the B200 engine (image inputs).

Same plugin contract as the reference's `Qwen2VLModel` (t2v_metrics/models/vqascore_models/qwen2vl_model.py:93-301): class
attributes, `forward(images, texts, question_template, answer_template, temperature) -> CPU fp32 Tensor[n]` with
score = softmax(last-position logits / temperature)[first answer token]. The reference loops over samples and calls
`generate(max_new_tokens=1)`; here the whole batch is ONE prefill in libvqa_b200.so and identical images are encoded once.
Video inputs (decord / qwen_vl_utils frame sampling) are outside this engine's scope and raise.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from ...config import QWEN25VL_MODELS as _TABLE, Qwen25VLConfig
from ...constants import HF_CACHE_DIR
from .qwen_utils import qwen_image_to_patches, build_prompt_ids, default_question_template, default_answer_template
from .vqa_model import VQAScoreModel

QWEN2_VL_MODELS: Dict[str, dict] = {name: dict(model=dict(path=spec["weights"]), config=spec["config"]) for name, spec in _TABLE.items()}


def _generation_config_penalty(checkpoint_path: str) -> float:
    """repetition_penalty of the generation_config.json stored beside the checkpoint (what `from_pretrained` would attach to
    `model.generation_config` in the reference, qwen2vl_model.py:116-130); 1.0 when there is none."""
    import json, os
    d = checkpoint_path if os.path.isdir(checkpoint_path) else os.path.dirname(checkpoint_path)
    f = os.path.join(d, "generation_config.json")
    if d and os.path.isfile(f):
        with open(f) as fh:
            return float(json.load(fh).get("repetition_penalty", 1.0) or 1.0)
    return 1.0


class Qwen2VLModel(VQAScoreModel):
    video_mode = "direct"
    allows_image = True
    supports_trace = False

    def __init__(self, model_name="qwen2.5-vl-7b", device="cuda", cache_dir=HF_CACHE_DIR, tokenizer=None,
                 state_dict: Optional[Dict[str, torch.Tensor]] = None, config: Optional[Qwen25VLConfig] = None,
                 checkpoint: Optional[str] = None, repetition_penalty: Optional[float] = None, **kwargs):
        assert model_name in QWEN2_VL_MODELS
        # The reference's scores come out of `generate(..., output_scores=True)`, i.e. AFTER the logits processors the checkpoint's
        # generation_config.json configures (Qwen2.5-VL-Instruct ships repetition_penalty 1.05). None = read it from the
        # generation_config.json next to `checkpoint` when there is one, else 1.0 (off).
        self.repetition_penalty = repetition_penalty
        self._tokenizer_override, self._state_dict, self._config_override, self._checkpoint = tokenizer, state_dict, config, checkpoint
        super().__init__(model_name=model_name, device=device, cache_dir=cache_dir)

    def load_model(self):
        from ...engine import QwenVLEngine
        spec = QWEN2_VL_MODELS[self.model_name]
        self.cfg: Qwen25VLConfig = self._config_override or spec["config"]()
        if self._tokenizer_override is not None:
            self.tokenizer = self._tokenizer_override
        else:
            from transformers import AutoTokenizer
            self.tokenizer = AutoTokenizer.from_pretrained(spec["model"]["path"], cache_dir=self.cache_dir)
        sd = self._state_dict
        if sd is None:
            path = self._checkpoint or spec["model"]["path"]
            import os
            if not os.path.isfile(path):
                raise FileNotFoundError(f"Qwen2.5-VL weights not found at {path!r}; pass `checkpoint=` or `state_dict=` (no network here)")
            if path.endswith(".safetensors"):
                from safetensors.torch import load_file
                sd = load_file(path)
            else:
                sd = torch.load(path, map_location="cpu")
        if self.repetition_penalty is None:
            self.repetition_penalty = _generation_config_penalty(self._checkpoint or spec["model"]["path"])
        dev = torch.device(self.device if self.device != "cuda" else "cuda:0")
        self.engine = QwenVLEngine(self.cfg, dev)
        self.engine.load_state_dict(sd)
        self._state_dict = None

    def load_images(self, image: List[str]):
        """-> (patches fp32 [sum P, 1176] on the device, [(1, gh, gw), ...])"""
        if any(p[-4:].lower() in (".mp4", ".avi", ".mov", ".mkv") for p in image):
            raise NotImplementedError("video inputs are outside the B200 engine's hot-path scope")
        # PIL decode on the host, everything else (smart_resize, PIL-exact bicubic, normalise, frame duplication, merge-order patch
        # rows) in ONE device kernel -- bit-identical to qwen_utils.qwen_image_to_patches, the CPU path the reference runs per image
        import numpy as np
        from ...engine import qwen_preprocess_u8
        raw = [torch.from_numpy(np.ascontiguousarray(np.asarray(self.image_loader(p).convert("RGB"), dtype=np.uint8))) for p in image]
        return qwen_preprocess_u8(raw, self.engine.device, self.cfg.patch_size, self.cfg.temporal_patch_size, self.cfg.spatial_merge_size)

    @torch.no_grad()
    def forward(self, images: List[str], texts: List[str], fps=None, question_template: str = default_question_template,
                answer_template: str = default_answer_template, max_new_tokens: int = 1, temperature: float = 1.0,
                debug: bool = False, repetition_penalty: Optional[float] = None) -> torch.Tensor:
        assert len(images) == len(texts), "Number of images/videos and texts must match"
        questions = [question_template.format(t) for t in texts]
        answers = [answer_template.format(t) for t in texts]
        uniq: Dict[str, int] = {}
        index = [uniq.setdefault(p, len(uniq)) for p in images]
        patches, grids = self.load_images(list(uniq.keys()))
        unit = self.cfg.spatial_merge_size ** 2
        prompts, answer_ids = [], []
        cache = self.__dict__.setdefault("_prompt_cache", {})
        if len(cache) > 65536:
            cache.clear()
        for q, a, img in zip(questions, answers, index):
            t, gh, gw = grids[img]
            prompts.append(build_prompt_ids(self.tokenizer, q, t * gh * gw // unit, self.cfg.image_token_id, cache))
            ids = cache.get(("answer", a))
            if ids is None:
                ids = cache[("answer", a)] = tuple(self.tokenizer.encode(a, add_special_tokens=False))
            if not ids:
                raise ValueError("empty answer")
            answer_ids.append(ids[0])      # max_new_tokens=1: only the first answer token is ever scored (qwen2vl_model.py:259-263)
        probs = self.engine.score_prompts(patches, grids, prompts, answer_ids, image_of_sample=index, temperature=temperature,
                                          repetition_penalty=self.repetition_penalty if repetition_penalty is None else repetition_penalty)
        return probs.float().cpu()
